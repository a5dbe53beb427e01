import torch


class IdentityMCObjective(torch.nn.Module):
    def forward(self, samples, X=None):
        return samples.squeeze(-1)


class IdentityMCMultiOutputObjective(torch.nn.Module):
    def forward(self, samples, X=None):
        return samples
