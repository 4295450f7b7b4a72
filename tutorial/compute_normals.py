"""Path-compatible entry point: ``cd tutorial && python compute_normals.py --input ... --mode ...``
(same flags as the reference's tutorial/compute_normals.py:17-28)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from deepfit_b200.compute_normals import main  # noqa: E402

if __name__ == '__main__':
    main()
