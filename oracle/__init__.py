"""CPU oracle for the tree-distance path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (treecl_b200) never does; it fails loudly when the CUDA library is missing.
"""
from .oracle import *  # noqa: F401,F403
from . import mds  # noqa: F401,E402   (classical MDS chain of treeCl/distance_matrix.py, NumPy)
