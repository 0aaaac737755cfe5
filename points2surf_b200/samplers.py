"""Index samplers of the reference's data loader (source/data_loader.py:71-176), same NumPy RandomState call sequence so
that seeded runs pick the same patches: SequentialPointcloudPatchSampler (`--sampling full`),
SequentialShapeRandomPointcloudPatchSampler (`--sampling sequential_shapes_random_patches`, and the
`random_shape_consecutive` training order), RandomPointcloudPatchSampler (`random` training order).
`data_source` needs `shape_names` and `shape_patch_count` (patches = query points per shape)."""
import numpy as np


class SequentialPointcloudPatchSampler:
    def __init__(self, data_source):
        self.data_source = data_source
        self.total_patch_count = int(sum(data_source.shape_patch_count[i] for i, _ in enumerate(data_source.shape_names)))

    def __iter__(self):
        return iter(range(self.total_patch_count))

    def __len__(self):
        return self.total_patch_count


class SequentialShapeRandomPointcloudPatchSampler:
    """All patches of a shape stay adjacent; inside a shape a random subset of at most `patches_per_shape` patches
    without replacement.  After iteration `shape_patch_inds[shape]` holds the chosen local indices (saved as
    `<shape>.idx` by the evaluation, source/points_to_surf_eval.py:292-294)."""

    def __init__(self, data_source, patches_per_shape, seed=None, sequential_shapes=False, identical_epochs=False):
        self.data_source = data_source
        self.patches_per_shape = patches_per_shape
        self.sequential_shapes = sequential_shapes
        self.seed = seed
        self.identical_epochs = identical_epochs
        self.shape_patch_inds = None
        if self.seed is None:
            self.seed = np.random.randint(0, 2**32 - 1)
        self.rng = np.random.RandomState(self.seed)
        self.total_patch_count = int(sum(min(self.patches_per_shape, data_source.shape_patch_count[i])
                                         for i, _ in enumerate(data_source.shape_names)))

    def __iter__(self):
        if self.identical_epochs:
            self.rng.seed(self.seed)
        counts = list(self.data_source.shape_patch_count)
        offsets = [0] + list(np.cumsum(counts))[:-1]
        shape_inds = range(len(self.data_source.shape_names))
        if not self.sequential_shapes:
            shape_inds = self.rng.permutation(shape_inds)
        self.shape_patch_inds = [[]] * len(self.data_source.shape_names)
        order = []
        for si in shape_inds:
            start, end = offsets[si], offsets[si] + counts[si]
            chosen = self.rng.choice(range(start, end), size=min(self.patches_per_shape, end - start), replace=False)
            order.extend(chosen)
            self.shape_patch_inds[si] = chosen - start
        return iter(order)

    def __len__(self):
        return self.total_patch_count


class RandomPointcloudPatchSampler:
    """A random subset of all patches of the data set, fully shuffled."""

    def __init__(self, data_source, patches_per_shape, seed=None, identical_epochs=False):
        self.data_source = data_source
        self.patches_per_shape = patches_per_shape
        self.seed = seed
        self.identical_epochs = identical_epochs
        if self.seed is None:
            self.seed = np.random.randint(0, 2**32 - 1)
        self.rng = np.random.RandomState(self.seed)
        self.total_patch_count = int(sum(min(self.patches_per_shape, data_source.shape_patch_count[i])
                                         for i, _ in enumerate(data_source.shape_names)))

    def __iter__(self):
        if self.identical_epochs:
            self.rng.seed(self.seed)
        return iter(self.rng.choice(sum(self.data_source.shape_patch_count), size=self.total_patch_count, replace=False))

    def __len__(self):
        return self.total_patch_count


def fixed_uniform_subsample_ids(num_points, sub_sample_size):
    """`--fixed_subsample 1` with `--uniform_subsample 1` (experiments/train_p2s_vanilla_uniform_subsample.sh):
    utils.get_point_cloud_sub_sample re-seeds its RandomState with 42 before every draw (source/base/utils.py:210-216), so every
    query of a shape gets the SAME ids, `RandomState(42).randint(0, N, S)` -- reproduced here bit for bit, stream included."""
    if num_points < sub_sample_size:
        raise ValueError('sub-sample needs N >= sub_sample_size (the reference zero-pads after an in-place shuffle; unsupported)')
    return np.random.RandomState(42).randint(low=0, high=num_points, size=sub_sample_size)
