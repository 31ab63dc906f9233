"""fa2_b200 -- Python host side of the B200-native FlashAttention-2 drop-in (ctypes over
libfa2_b200.so).  PyTorch is used only for device memory and streams."""
from ._lib import FA2Error, LIB_PATH, load  # noqa: F401
from .api import (SUPPORTED_HEAD_DIMS, backward, forward, forward_backward, partition, plan_split,  # noqa: F401
                  run_flash_attention, seq_range)
