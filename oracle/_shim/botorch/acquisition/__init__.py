from .monte_carlo import qExpectedImprovement, qNoisyExpectedImprovement, qUpperConfidenceBound  # noqa: F401
