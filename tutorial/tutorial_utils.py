"""Path-compatible shim for ``import tutorial_utils as tu`` (reference tutorial/compute_normals.py:13)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from deepfit_b200.tutorial_utils import *  # noqa: F401,F403,E402
from deepfit_b200.tutorial_utils import SinglePointCloudDataset, compute_principal_curvatures  # noqa: F401,E402
