"""B200-native drop-in for erweitern_55's batched policy/value network evaluation."""
import os as _os

# The host-buffer predict pipelines chunks over eight CUDA streams; with the default of 8 hardware queues per context
# they share queues with each other and with the context's main stream and late chunks wait behind early ones.
# Read by the driver when the CUDA context is created, so it must be in the environment before torch / libe55 touch CUDA.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
