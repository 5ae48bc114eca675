"""lmpc-b200: batched evaluation of a condensed linear-MPC law on B200.

Host-side mirror of the slice of LinearMPC.jl that sits on the hot path
(`compute_control`, `solve`, `Simulation`, `ExplicitMPC` evaluation), bound with ctypes to
the C ABI in include/lmpc_gpu.h.  The directory name contains a dot, so import it through
the repo-root shim:  ``import lmpc_b200``.

There is no CPU fallback: if liblmpcgpu.so is missing, or no sm_100 GPU is visible, every
call raises.
"""
from .host import (  # noqa: F401
    LmpcError, lib, lib_path, load_library, default_settings,
    GpuMPC, MultiGpuMPC, GpuExplicitMPC, ResultWindow, MPC, setup_gpu, host_ldp, fp64_peak_tflops, launch_count, device_count, pinned_empty,
    KERNEL_AUTO, KERNEL_THREAD, KERNEL_WARP, KERNEL_CTA,
)
from .sharding import shard_range, gather_columns, max_over_ranks  # noqa: F401,E402
from .examples import mpc_examples, zoh  # noqa: F401,E402
