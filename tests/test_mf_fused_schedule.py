"""Host-side statement of how the fused launches of the bulk-copy PDHG pass (csrc/pdhg_mf_bulk.cuh, FUSE) are scheduled,
and why the schedule is race-free: a numpy model of the launch -- blocks own contiguous runs of slabs, every block that holds
slabs of an instance recomputes that instance's small-vector update from what the PREVIOUS launch left (small state `src`,
partial sums of the read set), the block that owns the instance's first slab writes the new small state to `dst`, the passes
write the OTHER set of partial sums -- executed with the steps of the blocks interleaved at random.  Whatever the
interleaving, K launches + the closing update equal K steps of the sequential numpy statement (tests/mf_reference.MatrixFree,
itself proven equal to the CSR iteration on the oracle's matrix in tests/test_mf_reference.py).  With ONE set of partial sums
(what the first device version did) an interleaving exists that changes the result: the negative control at the bottom.
CPU only; the device kernels are compared with the same numpy statement in tests/test_pdhg_mf_gpu.py."""
import numpy as np
import pytest

from helpers import arrays_of
from mf_reference import EPS, MatrixFree
from neptune_mip_b200 import synth


class FusedLaunches:
    """B instances; global state per instance as the device holds it.  Small state sets q[0], q[1] and the canonical one;
    partial-sum sets P[0] (second set) and P[1] (canonical)."""

    def __init__(self, arrays, blocks, rng, single_p_set=False):
        self.mf = [MatrixFree(a) for a in arrays]
        self.B, self.F, self.N = len(arrays), arrays[0]["F"], arrays[0]["N"]
        self.blocks, self.rng, self.single = blocks, rng, single_p_set
        F, N = self.F, self.N
        z = np.zeros
        self.x = [z((F, N, N)) for _ in arrays]
        self.yS = [z((F, N, N)) for _ in arrays]
        small = lambda: dict(y1=z((F, N)), y2=z(N), y3=z((F, N)), y4=z(N), c=z((F, N)), cbar=z((F, N)))   # noqa: E731
        self.canon = [small() for _ in arrays]
        self.q = [[small() for _ in arrays] for _ in range(2)]
        pset = lambda: dict(P1=z((F, N)), P3=z((F, N)), P4=z((F, N)), PS=z((F, N)))                        # noqa: E731
        self.P = [[pset() for _ in arrays] for _ in range(2)]
        self.k = 0                                    # launches of the current chunk so far

    # ---- what one block does, as a list of steps (closures) in program order -------------------------------------
    def _small_update(self, b, src, pread, post, tau, sig):
        mf, a = self.mf[b], self.mf[b].a
        y1, y3, y4 = src["y1"], src["y3"], src["y4"]
        if post:
            s = sig * mf.S4
            v = y4 + s * pread["P4"].sum(axis=0)
            y4 = v - s * np.minimum(v / s, a["Kj"])
            s = sig * mf.S1
            v = y1 + s * (pread["P1"] - src["cbar"])
            y1 = v - s * np.maximum(v / s, -EPS)
            s = sig * mf.S3
            y3 = y3 + s * pread["P3"] - s * 1.0
        gc = -y1 + a["m"][:, None] * src["y2"][None, :] - pread["PS"]
        cn = np.clip(src["c"] - tau * mf.Tc * gc, 0.0, 1.0)
        cb = 2 * cn - src["c"]
        s = sig * mf.S2
        v = src["y2"] + s * (a["m"] @ cb)
        y2 = v - s * np.minimum(v / s, a["Mj"])
        return dict(y1=y1, y2=y2, y3=y3, y4=y4, c=cn, cbar=cb)

    def _block_steps(self, first, count, chunk_first, k):
        F = self.F
        rd, wr = (k + 1) & 1, k & 1
        if self.single:
            rd = wr = 1
        local = {}                                   # the block's shared-memory copy of the instance vectors

        def small(b, writer):
            def run():
                mf = self.mf[b]
                tau, sig = mf.eta / mf.omega, mf.eta * mf.omega
                src = self.canon[b] if chunk_first else self.q[(k + 1) & 1][b]
                new = self._small_update(b, src, self.P[rd][b], not chunk_first, tau, sig)
                local[b] = new
                if writer:
                    self.q[k & 1][b] = {key: val.copy() for key, val in new.items()}
            return run

        def tile(b, f):
            def run():
                mf, v = self.mf[b], local[b]
                tau, sig = mf.eta / mf.omega, mf.eta * mf.omega
                g = mf.obj[f] + v["y1"][f][None, :] + v["y3"][f][:, None] + mf.wr[f] * v["y4"][None, :] + self.yS[b][f]
                xn = np.clip(self.x[b][f] - tau * mf.Tx[f] * g, 0.0, 1.0)
                xb = 2 * xn - self.x[b][f]
                self.yS[b][f] = np.maximum(self.yS[b][f] + sig * mf.SS * (xb - v["cbar"][f][None, :]), 0.0)
                self.x[b][f] = xn
                p = self.P[wr][b]
                p["P1"][f] = xb.sum(axis=0); p["P3"][f] = xb.sum(axis=1)
                p["P4"][f] = (mf.wr[f] * xb).sum(axis=0); p["PS"][f] = self.yS[b][f].sum(axis=0)
            return run

        steps = []
        for n in range(count):
            slab = first + n
            b, f = divmod(slab, F)
            if n == 0 or f == 0:
                steps.append(small(b, f == 0))
            steps.append(tile(b, f))
        return steps

    def launch(self, chunk_first, order=None):
        total = self.B * self.F
        q, rem = divmod(total, self.blocks)
        queues = []
        for blk in range(self.blocks):
            first = blk * q + min(blk, rem)
            count = q + (1 if blk < rem else 0)
            queues.append(self._block_steps(first, count, chunk_first, self.k))
        # interleave: repeatedly pick a block that still has steps (random, or a forced order) and run its next step
        live = [blk for blk in range(self.blocks) if queues[blk]]
        if order is not None:                         # run the blocks one after the other in this order
            for blk in order:
                for step in queues[blk]:
                    step()
        else:
            while live:
                blk = live[self.rng.integers(len(live))]
                queues[blk].pop(0)()
                if not queues[blk]:
                    live.remove(blk)
        self.k += 1

    def close_chunk(self):
        """k_mf_small_from: POST of the last pass from the small state it left, everything back to the canonical arrays"""
        last = (self.k - 1) & 1
        for b in range(self.B):
            mf, a = self.mf[b], self.mf[b].a
            tau, sig = mf.eta / mf.omega, mf.eta * mf.omega
            src, p = self.q[last][b], self.P[1 if self.single else last][b]
            s = sig * mf.S4
            v = src["y4"] + s * p["P4"].sum(axis=0)
            y4 = v - s * np.minimum(v / s, a["Kj"])
            s = sig * mf.S1
            v = src["y1"] + s * (p["P1"] - src["cbar"])
            y1 = v - s * np.maximum(v / s, -EPS)
            s = sig * mf.S3
            y3 = src["y3"] + s * p["P3"] - s * 1.0
            self.canon[b] = dict(y1=y1, y2=src["y2"], y3=y3, y4=y4, c=src["c"], cbar=src["cbar"])
            # the solver refreshes the canonical column sums of yS before the next chunk (k_mf_eval, only_ps): the first
            # launch of a chunk reads them from the canonical set whatever set the last launch wrote
            self.P[1][b]["PS"] = self.yS[b].sum(axis=1)
        self.k = 0


def _instances(B, N, F, cores=40):
    return [arrays_of(synth.random_payload(N, F, s, node_cores=cores)) for s in range(B)]


def _sequential(arrays, iters):
    out = []
    for a in arrays:
        mf = MatrixFree(a)
        for _ in range(iters):
            mf.step()
        out.append(mf)
    return out


@pytest.mark.parametrize("B,N,F,blocks,iters,cores", [(3, 6, 4, 5, 7, 40), (2, 5, 3, 4, 4, 40), (4, 4, 2, 3, 5, 3), (1, 6, 5, 3, 6, 3),
                                                       (3, 6, 4, 7, 9, 3)])
def test_any_interleaving_of_the_blocks_equals_the_sequential_iteration(B, N, F, blocks, iters, cores):
    arrays = _instances(B, N, F, cores)
    want = _sequential(arrays, iters)
    for seed in range(4):
        sim = FusedLaunches(arrays, blocks, np.random.default_rng(seed))
        for k in range(iters):
            sim.launch(chunk_first=(k == 0))
        sim.close_chunk()
        for b in range(B):
            assert np.allclose(sim.x[b], want[b].x, rtol=0, atol=1e-13)
            assert np.allclose(sim.yS[b], want[b].yS, rtol=0, atol=1e-13)
            for key in ("y1", "y2", "y3", "y4", "c"):
                assert np.allclose(sim.canon[b][key], getattr(want[b], key), rtol=0, atol=1e-13), (key, b, seed)


def test_chunks_restart_from_the_canonical_arrays():
    """two chunks (3 + 4 launches): the second starts from the canonical small state the first one closed with"""
    arrays = _instances(2, 6, 3)
    want = _sequential(arrays, 7)
    sim = FusedLaunches(arrays, 4, np.random.default_rng(5))
    for chunk in (3, 4):
        for k in range(chunk):
            sim.launch(chunk_first=(k == 0))
        sim.close_chunk()
    for b in range(2):
        assert np.allclose(sim.x[b], want[b].x, rtol=0, atol=1e-13)
        assert np.allclose(sim.canon[b]["y4"], want[b].y4, rtol=0, atol=1e-13)
        assert np.allclose(sim.canon[b]["c"], want[b].c, rtol=0, atol=1e-13)


def test_one_set_of_partial_sums_is_not_enough():
    """negative control: with a single set of partial sums, a block that reaches an instance late reads sums the launch in
    progress has already overwritten (block 0 runs to its end before block 1 starts) -- the result changes.  This is the race
    the first device version had; tests/test_pdhg_mf_gpu.py::test_bulk_copy_pass_many_tiles_per_block caught it."""
    # one instance, two blocks (slabs 0-1 and 2-3), CPU rows that bind: the C4 dual sums the partial sums of ALL functions,
    # so it is the value a late block gets wrong
    arrays = _instances(1, 6, 4, cores=3)
    want = _sequential(arrays, 6)
    assert want[0].y4.max() > 0.0
    for order in ([0, 1], [1, 0]):
        sim = FusedLaunches(arrays, 2, np.random.default_rng(0), single_p_set=True)
        for k in range(6):
            sim.launch(chunk_first=(k == 0), order=order)
        sim.close_chunk()
        assert np.abs(sim.x[0] - want[0].x).max() > 1e-6, order
        ok = FusedLaunches(arrays, 2, np.random.default_rng(0))
        for k in range(6):
            ok.launch(chunk_first=(k == 0), order=order)
        ok.close_chunk()
        assert np.allclose(ok.x[0], want[0].x, rtol=0, atol=1e-13)
        assert np.allclose(ok.canon[0]["y4"], want[0].y4, rtol=0, atol=1e-13)
