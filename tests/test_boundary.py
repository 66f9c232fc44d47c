"""CPU: the C-ABI library loads, exports every symbol the header declares, the ctypes table covers
the header, and the product package never touches the oracle."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vidseg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vidseg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from vidseg_diffusion_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (vidseg_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.vidseg_abi_version() >= 1


def test_workspace_queries_without_gpu(lib):
    assert lib.vidseg_kmeans_workspace_bytes(14336, 640, 20, 10, 4) > 14336 * 640 * 4
    assert lib.vidseg_kmeans_workspace_bytes(0, 640, 20, 10, 4) == 0
    assert lib.vidseg_kmeans_workspace_bytes(100, 8, 5, 10, 9) == 0  # n_trials > 8
    assert lib.vidseg_refine_workspace_bytes(14, 1024, 640) >= 3 * 14 * 1024 * 640 * 4


def test_library_is_sm100a_with_no_other_arch():
    from vidseg_diffusion_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vidseg_diffusion_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "/root/reference" in txt:
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_host_api_fails_loudly_without_cuda():
    import pytest
    import torch
    from vidseg_diffusion_b200 import _lib
    from vidseg_diffusion_b200.features import aggregate_normalize
    from vidseg_diffusion_b200.kmeans import KMeans
    with pytest.raises(_lib.VidsegError):
        aggregate_normalize([torch.zeros(4, 3, 8)], 2)
    with pytest.raises(_lib.VidsegError):
        KMeans(n_clusters=2).fit_predict(torch.zeros(10, 4))


def test_kmeanspp_draws_match_sklearn_stream():
    """The host mirror must consume numpy's global stream exactly like sklearn's seeding."""
    import numpy as np
    from vidseg_diffusion_b200.kmeans import draw_kmeanspp_randoms, n_local_trials
    n, k, r = 1024, 5, 10
    np.random.seed(3)
    first, rand = draw_kmeanspp_randoms(n, k, r)
    pos = np.random.get_state()[2]
    np.random.seed(3)
    rs = np.random.mtrand._rand
    sw = np.ones(n, dtype=np.float32)
    for i in range(r):
        assert rs.choice(n, p=sw / sw.sum()) == first[i]
        for c in range(k - 1):
            assert np.array_equal(rs.uniform(size=n_local_trials(k)), rand[i, c])
    assert np.random.get_state()[2] == pos
    assert rand.shape == (r, k - 1, 3) and n_local_trials(20) == 4 and n_local_trials(50) == 5
