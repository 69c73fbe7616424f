"""Input side of the hot path: the reference's meta-file datasets
(datasets/example_dataset.py, target_dataset.py, example_loader.py) with the per-image pixel work moved to
the device (csrc/image_ops.cu)."""
from .example_dataset import (ExampleDataset, ExampleTransform, SyntheticPairs, TargetDataset, collate,  # noqa: F401
                              parse_meta_file, prepare_image)
