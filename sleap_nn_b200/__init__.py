"""sleap_nn_b200 - B200-native (sm_100a) implementation of sleap-nn's heat-map hot path.

Drop-in module layout (same names, signatures and return conventions as the reference):

    sleap_nn_b200.inference.peak_finding      <-> sleap_nn.inference.peak_finding
    sleap_nn_b200.inference.paf_grouping      <-> sleap_nn.inference.paf_grouping
    sleap_nn_b200.inference.ops.{peaks,crops,paf}
    sleap_nn_b200.inference.utils             (interp1d)
    sleap_nn_b200.data.{confidence_maps,edge_maps,utils,instance_cropping}

Every function routes into hand-written CUDA kernels through the C ABI in
include/sleapnn_b200.h (ctypes, `_native.py`).  There is no CPU / PyTorch fallback.
`sleap_nn_b200.compat.install()` aliases these modules over the reference's import paths.
"""

__version__ = "0.1.0"
