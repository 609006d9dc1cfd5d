"""(a) assembly parity: the device CSR must EQUAL the oracle's (== the reference's own) matrix."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import KIND_NAMES, arrays_of, cuda_batch, small_payloads
from neptune_mip_b200 import synth
from oracle import model as omodel

pytestmark = pytest.mark.gpu


def _compare(mdl, b, ref):
    A = ref["A"]
    assert mdl.rows == A.shape[0] and mdl.cols == A.shape[1] and mdl.nnz == A.nnz
    assert np.array_equal(mdl.row_ptr.cpu().numpy(), A.indptr)
    assert np.array_equal(mdl.col_idx.cpu().numpy(), A.indices)
    assert np.array_equal(mdl.val[b].cpu().numpy(), A.data)
    for k in ("obj", "lo", "hi"):
        assert np.array_equal(getattr(mdl, k)[b].cpu().numpy(), ref[k]), k
    assert np.array_equal(mdl.col_lb[b].cpu().numpy(), ref["lb"])
    assert np.array_equal(mdl.col_ub[b].cpu().numpy(), ref["ub"])
    assert np.array_equal(mdl.col_int.cpu().numpy(), ref["integ"])
    At = omodel.transpose_csr(A)
    assert np.array_equal(mdl.rowT_ptr.cpu().numpy(), At.indptr)
    assert np.array_equal(mdl.colT_idx.cpu().numpy(), At.indices)
    assert np.array_equal(mdl.valT[b].cpu().numpy(), At.data)


@pytest.mark.parametrize("name,payload,alpha", small_payloads(), ids=lambda v: v if isinstance(v, str) else None)
@pytest.mark.parametrize("kind", KIND_NAMES)
def test_assemble_equals_oracle(name, payload, alpha, kind):
    from neptune_mip_b200 import device
    a = arrays_of(payload)
    mdl = device.assemble(cuda_batch([payload]), kind, alpha)
    _compare(mdl, 0, omodel.build_step1(a, kind, alpha))


@pytest.mark.parametrize("kind", KIND_NAMES)
def test_assemble_batched(kind):
    from neptune_mip_b200 import device
    payloads = [synth.random_payload(20, 5, s, node_cores=100) for s in range(5)]
    mdl = device.assemble(cuda_batch(payloads), kind, 0.5)
    for b, p in enumerate(payloads):
        _compare(mdl, b, omodel.build_step1(arrays_of(p), kind, 0.5))


def test_assemble_c2_and_sizes():
    from neptune_mip_b200 import device
    p = synth.config_payload("C2")
    mdl = device.assemble(cuda_batch([p]), "min_delay")
    assert (mdl.cols, mdl.rows, mdl.nnz) == (25500, 1600, 101500)
    _compare(mdl, 0, omodel.build_step1(arrays_of(p), "min_delay"))
    assert device.model_sizes(500, 50, 0) == (76000, 12525000, 50075000)
    assert device.model_sizes(2000, 200, 0) == (1204000, 800400000, 3201200000)


def test_strengthened_rows_are_appended_after_reference_rows():
    from neptune_mip_b200 import device
    from neptune_mip_b200._lib import FLAG_STRENGTHEN
    p = synth.random_payload(8, 4, 0, node_cores=30)
    a = arrays_of(p)
    N, F = a["N"], a["F"]
    base = omodel.build_step1(a, "min_delay_util", 0.5)
    mdl = device.assemble(cuda_batch([p]), "min_delay_util", 0.5, flags=FLAG_STRENGTHEN)
    A = sp.csr_matrix((mdl.val[0].cpu().numpy(), mdl.col_idx.cpu().numpy(), mdl.row_ptr.cpu().numpy()),
                      shape=(mdl.rows, mdl.cols))
    R = base["A"].shape[0]
    top = A[:R]
    assert np.array_equal(top.indptr, base["A"].indptr) and np.array_equal(top.indices, base["A"].indices)
    assert np.array_equal(top.data, base["A"].data)
    X = F * N * N
    S = A[R:].toarray()
    assert S.shape[0] == X
    for q in range(X):
        f, j = q // (N * N), q % N
        row = np.zeros(mdl.cols); row[q] = 1.0; row[X + f * N + j] = -1.0
        assert np.array_equal(S[q], row)
    assert np.all(mdl.hi[0, R:].cpu().numpy() == 0.0) and np.all(np.isneginf(mdl.lo[0, R:].cpu().numpy()))
    # reference rows keep their bounds, except C1a (big-M linking, implied by S) which is left free
    lo, hi = mdl.lo[0, :R].cpu().numpy(), mdl.hi[0, :R].cpu().numpy()
    c1a = np.zeros(R, dtype=bool); c1a[0:2 * F * N:2] = True
    assert np.array_equal(lo, base["lo"]) and np.array_equal(hi[~c1a], base["hi"][~c1a])
    assert np.all(np.isposinf(hi[c1a]))
    At = omodel.transpose_csr(A)
    assert np.array_equal(mdl.rowT_ptr.cpu().numpy(), At.indptr)
    assert np.array_equal(mdl.colT_idx.cpu().numpy(), At.indices)
    assert np.array_equal(mdl.valT[0].cpu().numpy(), At.data)
