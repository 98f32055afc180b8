"""CPU suite: the C-ABI library loads, exports every symbol include/vo_b200.h declares, and fails
loudly (no fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vo_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"libvo_b200.so does not export {n}"
    assert sorted(_lib.EXPORTS) == names
    assert L.vo_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vo_b200
    with pytest.raises(vo_b200.VoError, match="no CPU fallback"):
        vo_b200.Context(0)


def test_null_context_is_rejected_by_the_state_entry_points():
    """Entry points that only touch context state answer a null context with an error, not a crash."""
    import ctypes as C
    from vo_b200 import _lib
    L = _lib.lib()
    assert L.vo_frames_use_graph(None, 1) != 0 and b"null" in L.vo_last_error()
    assert L.vo_frames_graph_state(None) == 0
    stats = (C.c_int * 4)()
    assert L.vo_match_stats(None, stats) != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "r7020e-visual-odometry_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "vo_oracle" not in txt, f


def test_mex_gateways_export_mexfunction():
    d = os.path.join(ROOT, "r7020e-visual-odometry_b200", "csrc", "mex")
    for name in ("vo_sift_mex", "vo_match_mex", "vo_triangulate_mex", "vo_p3p_mex", "vo_frames_mex"):
        so = os.path.join(d, name + ".mexa64")
        assert os.path.exists(so), f"{so} missing: run python __graft_entry__.py"
        lib = C.CDLL(so, mode=os.RTLD_LAZY)
        assert hasattr(lib, "mexFunction")


def test_gateways_fail_loudly_without_a_gpu_and_check_their_arguments():
    """The MEX gateways through the stand-in MATLAB host on a CPU-only box: argument errors are raised before any
    device work, and a valid call ends in mexErrMsgIdAndTxt("vo:ctx:create", ... no CPU fallback), never in a result."""
    import numpy as np
    import torch
    from vo_b200 import mexhost, reloc, shard, io   # noqa: F401  (the host-side modules import without a GPU)
    h = mexhost.Host()
    f = np.zeros((4, 128), dtype=np.float32)
    with pytest.raises(mexhost.MexError, match="vo:match:class"):
        h.call("vo_match_mex", 1, f.astype(np.float64), f.astype(np.float64))
    with pytest.raises(mexhost.MexError, match="vo:frames:class"):
        h.call("vo_frames_mex", 1, np.zeros((8, 8, 2), np.float32), np.zeros((8, 8, 2), np.float32), np.eye(3, 4), np.eye(3, 4))
    with pytest.raises(mexhost.MexError, match="vo:frames:size"):
        h.call("vo_frames_mex", 1, np.zeros((8, 8, 2), np.uint8), np.zeros((8, 8, 3), np.uint8), np.eye(3, 4), np.eye(3, 4))
    with pytest.raises(mexhost.MexError, match="vo:sift:nargin"):
        h.call("vo_sift_mex", 1)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mexhost.MexError, match="no CPU fallback"):
        h.call("vo_match_mex", 1, f, f)
    with pytest.raises(mexhost.MexError, match="no CPU fallback"):
        h.call("vo_frames_mex", 1, np.zeros((8, 8, 2), np.uint8), np.zeros((8, 8, 2), np.uint8), np.eye(3, 4), np.eye(3, 4))
