"""Mirror of the hot-path slice of the reference's ``parallelism`` package (re-exports as parallelism/__init__.py:1-112)."""
from .communication import (all_gather, all_reduce, barrier, broadcast, gather_along_sequence_dim, get_rank,  # noqa: F401
                            get_world_size, initialize_distributed, reduce_scatter, ring_exchange, scatter,
                            scatter_along_sequence_dim, setup_sequence_parallel_group)
from .parallel_utils import (divide, ensure_divisibility, gather_tensor_along_dim, get_partition_start_end,  # noqa: F401
                             get_tensor_model_parallel_group, get_tensor_model_parallel_rank,
                             get_tensor_model_parallel_world_size, initialize_tensor_parallel, is_power_of_two,
                             split_tensor_along_dim)
from .ring import ring_attention_forward  # noqa: F401
from .sequence_parallel import (SequenceParallelAttention, SequenceParallelConfig, SequenceParallelConverter,  # noqa: F401
                                SequenceParallelMLP, SequenceShardedModule)
from .tensor_parallel import (ColumnParallelLinear, ModelParallelConverter, RowParallelLinear, TensorParallelAttention,  # noqa: F401
                              TensorParallelConfig, TensorParallelMLP)
