from .samplers import IIDNormalSampler, SobolQMCNormalSampler  # noqa: F401
