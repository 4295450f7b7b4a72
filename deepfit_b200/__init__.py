"""deepfit_b200 -- B200-native (sm_100a) implementation of DeepFit's inference hot path.

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/deepfit_b200.h``); ``_lib`` ctypes binding;
``ops`` one torch-facing wrapper per stage; ``DeepFit`` / ``tutorial_utils`` / ``compute_normals``
drop-ins for the reference's modules of the same name; ``pipeline`` the whole-cloud loop and the
multi-GPU sharding helpers.  Nothing here imports ``oracle/`` and nothing falls back to the CPU.
"""
__version__ = "0.1.0"
