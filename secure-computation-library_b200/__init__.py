"""B200 (sm_100a) implementation of SCL's data-parallel hot path.

Only what that path needs lives here:

* ``csrc/``      -- the hand-written CUDA kernels and the C ABI (libsclgpu.so,
                    declared in include/sclgpu.h);
* ``binding``    -- ctypes declarations of that ABI;
* ``api``        -- ``Context``: a thin host handle used by the tests / bench
                    (numpy host buffers or torch device tensors in, same out);
* ``sharding``   -- batch-of-secrets partitioning over the GPUs of one box.

The directory name contains a hyphen, so it is loaded through
``__graft_entry__.load_package()`` under the module name ``scl_b200``.
There is no CPU fallback anywhere in this package.
"""
from . import api, binding, sharding  # noqa: F401
from .api import Context, CudaError, InvalidArgument, LogicError, MultiContext  # noqa: F401
