"""torchrun entry: function-block-sharded PDHG vs the unsharded iteration (parity) and timing.
  python -m torch.distributed.run --nproc-per-node 2 tools/sharded_run.py parity|solve|time|mf-parity|mf-solve|mf-time N F [iters]
(the mf-* modes drive the matrix-free sharded solver, neptune_mip_b200/sharded_mf.py)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from neptune_mip_b200 import synth, sharding
from neptune_mip_b200.core.utils import data_to_solver_input
from neptune_mip_b200.sharded import ShardedLP

mode, N, F = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 100
backend = os.environ.get("NEPTUNE_DIST_BACKEND", "nccl")
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
ngpu = torch.cuda.device_count()
torch.cuda.set_device(local % ngpu)
dist.init_process_group(backend, **({"device_id": torch.device("cuda", local)} if backend == "nccl" else {}))
_nc = os.environ.get("NEPTUNE_NODE_CORES")
data = data_to_solver_input(synth.random_payload(N, F, 1 if _nc else 0, node_cores=int(_nc) if _nc else None), 1, with_db=False)
if mode.startswith("mf-"):
    from neptune_mip_b200.sharded_mf import ShardedMF
    lp = ShardedMF(data)
    if mode == "mf-parity":
        lp.iterate(iters)
        xloc = lp.x[0, : lp.Fg * N * N].reshape(lp.Fg, N * N).contiguous()
        xall = sharding.gather_rows(xloc.cpu(), F) if backend != "nccl" else sharding.gather_rows(xloc, F).cpu()
        ycoup = torch.cat([lp.y[0, lp.r2:lp.r2 + N], lp.y[0, lp.r4:lp.r4 + N]]).cpu()
        kkt = lp._kkt(lp.x, lp.y)                    # collective: every rank calls it
        if rank == 0:
            import neptune_mip_b200.sharded_mf as shm
            shm.world = lambda: (0, 1)               # reference: every function on one rank (no exchange)
            ref = ShardedMF(data)
            ref.iterate(iters)
            xr = ref.x[0, : F * N * N].reshape(F, N * N).cpu()
            yr = torch.cat([ref.y[0, ref.r2:ref.r2 + N], ref.y[0, ref.r4:ref.r4 + N]]).cpu()
            kr = ref._kkt(ref.x, ref.y)
            dx = float((xall - xr).abs().max()); dy = float((ycoup - yr).abs().max())
            print(f"MF-PARITY world={ws} N={N} F={F} iters={iters} max|dx|={dx:.3e} max|dy_coupling|={dy:.3e} "
                  f"kkt sharded={kkt} single={kr} exchanged doubles/iter={lp.exchanged_doubles / max(lp.iters, 1):.1f}", flush=True)
            assert dx <= 1e-9 * (1 + float(xr.abs().max())) and dy <= 1e-9 * (1 + float(yr.abs().max()))
            assert all(abs(u - v) <= 1e-9 * (1 + abs(v)) for u, v in zip(kkt, kr))
    elif mode == "mf-solve":
        info = lp.solve(max_iters=iters, check_every=128, eps_rel=1e-6)
        if rank == 0:
            from neptune_mip_b200 import device
            _, _, res = device.pdhg_mf_solve(device.InstanceBatch.from_datas([data]), max_iters=iters, check_every=128,
                                             eps_rel=1e-6, eps_abs=1e-9)
            print(f"MF-SOLVE world={ws} N={N} F={F} sharded: {info} | single GPU: primal={res[0]['primal_obj']:.9g} "
                  f"dual={res[0]['dual_obj']:.9g} iters={res[0]['iters']} converged={res[0]['converged']}", flush=True)
            tol = 1e-4 * (1 + abs(float(res[0]["primal_obj"])))
            assert info["converged"] and res[0]["converged"] == 1 and abs(info["primal_obj"] - float(res[0]["primal_obj"])) <= tol
    else:
        lp.iterate(5)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lp.iterate(iters); e1.record(); e1.synchronize()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), device="cuda")
        by = 64 * lp.X + 112 * lp.Cn + 8 * N * N
        if rank == 0:
            print(f"MF-TIME world={ws} N={N} F={F} (Fg={lp.Fg}) {ms / iters * 1e3:.1f} us/iter  {by * iters / ms / 1e6:.1f} GB/s per GPU  "
                  f"exchange {16 * N} B/iter", flush=True)
    dist.destroy_process_group()
    sys.exit(0)
lp = ShardedLP(data)
if mode == "parity":
    lp.iterate(iters)
    obj = lp.primal_objective()
    # gather x blocks (x columns of this rank's functions) and the replicated coupling multipliers
    xloc = lp.x[0, : lp.Fg * N * N].reshape(lp.Fg, N * N).contiguous()
    if backend != "nccl":
        xall = sharding.gather_rows(xloc.cpu(), F)
    else:
        xall = sharding.gather_rows(xloc, F).cpu()
    ycoup = lp._coupling(lp.y).cpu()
    if rank == 0:
        # reference: the same iteration with every function on one rank (no exchange)
        import neptune_mip_b200.sharded as sh
        sh.world = lambda: (0, 1)
        ref = ShardedLP(data)
        ref.iterate(iters)
        xr = ref.x[0, : F * N * N].reshape(F, N * N).cpu()
        yr = ref._coupling(ref.y).cpu()
        robj = float((ref.model.obj * ref.x).sum())
        dx = float((xall - xr).abs().max()); dy = float((ycoup - yr).abs().max())
        print(f"PARITY world={ws} N={N} F={F} iters={iters} max|dx|={dx:.3e} max|dy_coupling|={dy:.3e} "
              f"obj sharded={obj:.9g} single={robj:.9g}", flush=True)
        assert dx <= 1e-8 * (1 + float(xr.abs().max())) and dy <= 1e-8 * (1 + float(yr.abs().max()))
        assert abs(obj - robj) <= 1e-9 * (1 + abs(robj))
elif mode == "solve":
    # restarted solve to the LP optimum; rank 0 compares with the single-GPU solver on the same model
    info = lp.solve(max_iters=iters, check_every=128, eps_rel=1e-6)
    if rank == 0:
        from neptune_mip_b200 import device
        full = device.assemble(device.InstanceBatch.from_datas([data]), "min_delay")
        _, _, res = device.pdhg_solve(full, max_iters=iters, eps_rel=1e-6, eps_abs=1e-9)
        print(f"SOLVE world={ws} N={N} F={F} sharded: {info} | single GPU: primal={res[0]['primal_obj']:.9g} "
              f"dual={res[0]['dual_obj']:.9g} iters={res[0]['iters']} converged={res[0]['converged']}", flush=True)
        # the optimum lies between the single-GPU solver's dual and primal values (it may not have converged)
        lo_b, hi_b = float(res[0]["dual_obj"]), float(res[0]["primal_obj"])
        tol = 1e-4 * (1 + abs(hi_b))
        assert info["converged"] and lo_b - tol <= info["primal_obj"] <= hi_b + tol
else:
    lp.iterate(5)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lp.iterate(iters); e1.record(); e1.synchronize()
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    by, ex = lp.bytes_per_iteration()
    obj_avg = lp.primal_objective(True)          # collective: every rank must call it
    if rank == 0:
        m = lp.model
        print(f"TIME world={ws} N={N} F={F} (Fg={lp.Fg}) local rows={m.rows} cols={m.cols} nnz={m.nnz} "
              f"{ms / iters * 1e3:.1f} us/iter  {by * iters / ms / 1e6:.1f} GB/s per GPU  exchange {ex} B/iter "
              f"obj={obj_avg:.6g}", flush=True)
dist.destroy_process_group()
