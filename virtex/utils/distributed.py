from virtex_b200.distributed import (init_from_env, get_world_size, get_rank, is_master_process, synchronize,  # noqa: F401
                                     average_across_processes)
