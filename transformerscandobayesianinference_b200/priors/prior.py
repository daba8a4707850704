from torch.utils.data import DataLoader


class PriorDataLoader(DataLoader):
    """Protocol of a prior data loader (reference priors/prior.py): `__init__(num_steps, ...)`, attributes
    `num_features`, `num_outputs`, `fuse_x_y`, optionally `validate(model)`."""
    pass
