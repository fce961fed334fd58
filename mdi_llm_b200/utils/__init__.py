"""Support utilities (checkpoint I/O, converters, data loading, plotting, misc helpers)."""
from .checkpoint import (incremental_save, init_from_state_dict, lazy_load, load_from_pt, load_sd,  # noqa: F401
                         save_config, write_random_checkpoint)
from .context_managers import catch_loop_errors  # noqa: F401
from .data_loader import get_batch, load_dataset, split_dataset  # noqa: F401
from .misc import (detect_stop_tokens, estimate_loss, find_eot, get_lr, get_obj_size, loading_bar,  # noqa: F401
                   remove_prefix, waiting_animation)
from .plots import plot_tokens_per_time  # noqa: F401
from ..models.partition import count_transformer_blocks, split_and_store, split_parameters  # noqa: F401
