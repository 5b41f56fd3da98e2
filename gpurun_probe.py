import json, numpy as np, sys
sys.path.insert(0, '.')
import vectorizedadjoint_b200 as va, oracle
g = json.load(open('tests/golden/reference_goldens.json'))
mu0 = 1e3
x0 = [2.0, -2.0 / 3.0 + 10.0 / (81.0 * mu0) - 292.0 / (2187.0 * mu0 * mu0)]
for tol in ["1e-3", "1e-4", "1e-5", "1e-6", "1e-7", "1e-8", "1e-9", "1e-10", "1e-12"]:
    gg = g["ref17"][f"vanderpol_rkf78_{tol}"]
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_RKF78, True, float(tol), float(tol), n_out=2, max_steps=2048) as e:
        r = e.forward_adjoint([x0], [[mu0]], 0.0, 0.5, 1e-3, objective=va.OBJ_SEED, seeds=[[[1, 0], [0, 1]]])
    ex = np.abs(r["x_final"][0]-gg["x_final"]).max()/np.abs(gg["x_final"]).max()
    el = np.abs(r["lam"][0]-gg["lam"]).max()/np.abs(gg["lam"]).max()
    em = np.abs(r["mu"][0]-gg["mu"]).max()/np.abs(gg["mu"]).max()
    print(tol, r["n_accept"][0], gg["steps"], r["n_reject"][0], f"x {ex:.2e} lam {el:.2e} mu {em:.2e}", r["mu"][0].ravel())
