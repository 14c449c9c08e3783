"""Drop-in overlay for src/lib/models/data_parallel.py (`from models.data_parallel import DataParallel`,
trains/base_trainer.py:8,31-35).  Launched as one process per GPU (torchrun), the same call returns a sharded
replica whose backward all-reduces the gradients over NCCL; in a single process it is torch.nn.DataParallel."""
from centernet_b200.data_parallel import DataParallel, ShardedDataParallel, GradientAllReducer  # noqa: F401
