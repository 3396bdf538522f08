"""Diagnostic (GPU box): parity statistics of the CUDA path against the fp32 and fp64 oracles."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle
from tfmpc_b200 import envs, ops
from tfmpc_b200.envs import synthetic
from tfmpc_b200.solvers.ilqr import iLQR

def case(cfg, B, T, seed):
    rng = np.random.RandomState(seed)
    x0 = synthetic.sample_x0(cfg, B, rng)
    c = cfg["config"]
    if cfg["cls_name"] == "Navigation": lo, hi = np.ravel(c["low"]), np.ravel(c["high"])
    else: lo, hi = np.zeros(x0.shape[1]), np.ones(x0.shape[1])
    return x0, synthetic.sample_u_init(lo, hi, B, T, rng)

def gpu(cfg, prec, x0, u0):
    dt = torch.float32 if prec == "f32" else torch.float64
    e = envs.make_env(cfg); e.dtype = dt
    out = iLQR(e, dtype=dt).solve_device(x0, u0.shape[1], u_init=u0)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}

def cmp(tag, ia, ca, ib, cb):
    d = np.abs(ia - ib); relc = np.abs(ca - cb) / np.abs(cb)
    print(f"{tag:34s} same={np.mean(d==0):.4f} |d|<=1={np.mean(d<=1):.4f} |d|<=3={np.mean(d<=3):.4f} max|d|={d.max():3d} "
          f"relcost: med={np.median(relc):.1e} p99={np.percentile(relc,99):.1e} max={relc.max():.1e} frac>1e-4={np.mean(relc>1e-4):.4f} "
          f"[same-it] max={relc[d==0].max():.1e}")

o32, o64 = oracle.Oracle("f32"), oracle.Oracle("f64")
for name, cfg, B, T in [("nav_h50", synthetic.navigation_config(), 4096, 50), ("hvac6", synthetic.hvac_grid_config(2, 3), 256, 48),
                        ("hvac32", synthetic.hvac_grid_config(4, 8), 48, 48), ("res4", synthetic.reservoir_config(4), 256, 40),
                        ("res20", synthetic.reservoir_config(20), 48, 40)]:
    x0, u0 = case(cfg, B, T, 11)
    g32, g64 = gpu(cfg, "f32", x0, u0), gpu(cfg, "f64", x0, u0)
    r32 = o32.ilqr_solve(o32.make_env(cfg), x0, u0); r64 = o64.ilqr_solve(o64.make_env(cfg), x0, u0)
    print(f"== {name} B={B}: mean its gpu32={g32['stats'][:,0].mean()+1:.2f} orc32={r32['iterations'].mean()+1:.2f} orc64={r64['iterations'].mean()+1:.2f}"
          f" status gpu32={np.bincount(g32['stats'][:,3],minlength=5)} orc32={np.bincount(r32['status'],minlength=5)} gpu64={np.bincount(g64['stats'][:,3],minlength=5)} orc64={np.bincount(r64['status'],minlength=5)}")
    cmp("gpu32 vs orc32", g32["stats"][:, 0], g32["costs"].sum(1), r32["iterations"], r32["costs"].sum(1))
    cmp("gpu32 vs orc64", g32["stats"][:, 0], g32["costs"].sum(1), r64["iterations"], r64["costs"].sum(1))
    cmp("orc32 vs orc64 (fp32 noise band)", r32["iterations"], r32["costs"].sum(1), r64["iterations"], r64["costs"].sum(1))
    cmp("gpu64 vs orc64", g64["stats"][:, 0], g64["costs"].sum(1), r64["iterations"], r64["costs"].sum(1))

# box-QP
rng = np.random.RandomState(3)
for prec, o in (("f32", o32), ("f64", o64)):
    dt = torch.float32 if prec == "f32" else torch.float64
    for m in (2, 3, 6):
        B = 2000
        A = rng.normal(size=(B, m, m)); H = A @ np.swapaxes(A, 1, 2) + 0.3 * np.eye(m)
        q = rng.normal(size=(B, m)) * 3; lo, hi = -rng.uniform(0.05, 1.5, size=(B, m)), rng.uniform(0.05, 1.5, size=(B, m)); x0 = (lo + hi) / 2
        cu = lambda a: torch.as_tensor(a).to("cuda", dt).contiguous()
        out = ops.boxqp(cu(H), cu(q), cu(lo), cu(hi), cu(x0)); r = o.boxqp(H, q, lo, hi, x0)
        same = (out["free"].cpu().numpy() == r["free"]).all(1)
        print(f"boxqp {prec} m={m}: same free set {same.mean():.4f}; max |dx| all={np.abs(out['x'].cpu().numpy()-r['x']).max():.2e} same={np.abs(out['x'].cpu().numpy()-r['x'])[same].max():.2e}")
# LQR batched f64
rng = np.random.RandomState(0); B, T = 4096, 10
goal = rng.uniform(-10, 10, size=(B, 2)); x0 = rng.normal(size=(B, 2))
lq = envs.make_lqr_linear_navigation(goal, 5.0); lq.__init__(lq.F, lq.f, lq.C, lq.c, dtype=torch.float64)
out = lq.solve_device(x0, T, want_policy=True, want_value=True)
F = np.concatenate([np.eye(2), np.eye(2)], axis=1); c = np.concatenate([-2 * goal, np.zeros_like(goal)], axis=1)
r = o64.lqr_solve(F, np.zeros(2), np.diag([2.0, 2.0, 10.0, 10.0]), c, x0, T)
for key in ("states", "actions", "costs", "K", "k", "V", "v", "const"):
    a, b = out[key].cpu().numpy(), r[key]
    print("lqr f64", key, np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))
