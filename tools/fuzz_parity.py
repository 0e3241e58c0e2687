"""Randomised parity sweep on the GPU (not part of the test suite; test infrastructure, it imports the oracle): random shapes around the
block / tile / strip boundaries, every flag combination of rc_lml_grad, predict, and Sobol subsets, against oracle/ on identical inputs.
    python tools/fuzz_parity.py [seconds] [seed]"""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
from romcomma import _capi as C, gf_compat as gf
from oracle import gp, sobol
from conftest import random_problem, assert_close

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
edges = [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 256, 257, 300, 383, 384, 385, 511, 512, 513, 640, 700]
t0, cases = time.time(), 0
while time.time() - t0 < budget:
    L = int(rng.integers(1, 5))
    N = int(rng.choice(edges)) if rng.random() < 0.7 else int(rng.integers(1, 700))
    N = max(N, 2)                                   # Y is standardised per column
    if L * N > 1500:
        N = 1500 // L
    M = int(rng.integers(1, 13)) if rng.random() < 0.75 else int(rng.integers(13, 21))       # 13..20: the two-CTA instantiations of the sweep kernels
    full_F = bool(rng.random() < 0.5) and L > 1
    X, Y, ls, F, E = random_problem(N, M, L, seed=int(rng.integers(1 << 30)), full_F=full_F)
    dX, dY = C.dev(X), C.dev(Y)
    args = (C.dev(ls), C.dev(F[None]), C.dev(E[None]))
    ref = gp.lml_grad_mo(X, Y, ls, F, E)
    tag = f'N={N} M={M} L={L} full_F={full_F}'
    for flags in (C.RC_GRAD_NONE, C.RC_GRAD_VARIANCE, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES) + ((C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL,) if not full_F else ()):
        plan = C.LmlGradPlan(dX, dY, L, 1, flags)
        res = plan.unpack(plan(*args).cpu().numpy())[0]
        assert plan.info.cpu().tolist() == [0], tag
        assert_close(res['lml'], ref['lml'], what=f'{tag} flags={flags} lml')
        if flags & C.RC_GRAD_VARIANCE:
            dF, rF = (np.diag(res['dF']), np.diag(ref['dF'])) if flags & C.RC_GRAD_F_DIAGONAL else (res['dF'], ref['dF'])
            assert_close(dF, rF, atol=1e-10 * L * N, what=f'{tag} flags={flags} dF')
            assert_close(res['dE'], ref['dE'], atol=1e-10 * L * N, what=f'{tag} flags={flags} dE')
        if flags & C.RC_GRAD_LENGTHSCALES:
            assert_close(res['dls'], ref['dls'], atol=1e-10 * L * N, what=f'{tag} dls')
    ns = int(rng.integers(1, 40))
    xs = rng.normal(size=(ns, M))
    mean, var = gf.predict_core(dX, dY, args[0], F[None], E[None], C.dev(xs), L, 1, True)
    rm, rv = gp.predict_mo(X, Y, ls, F, E, xs)
    assert_close(mean[0].cpu().numpy(), rm, rtol=1e-7, atol=1e-9, what=f'{tag} predict mean')
    assert_close(var[0].cpu().numpy(), rv, rtol=1e-7, atol=1e-9, what=f'{tag} predict var')
    if not full_F:
        Fd = np.diag(F).copy()
        KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
        Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(ls), C.dev(Fd), C.dev(KiY.reshape(L, N)), True)
        masks = [int(rng.integers(1, 1 << M)) for _ in range(int(rng.integers(1, 4)))] + [C.slice_mask(0, M), C.slice_mask(0, 1), C.slice_mask(M - 1, M)]
        V = C.sobol_contract(dX, Phi, g0KY, L, True, masks).cpu().numpy()
        scale = np.abs(V).max() + 1e-300
        for v, m in zip(V, masks):
            idx = [i for i in range(M) if (m >> i) & 1]
            # V is a sum of N^2 terms of either sign: the comparison can only be as tight as the conditioning of the sum allows - the float64
            # oracle itself is 8 eps sum|terms| away from an extended-precision evaluation in the worst cases this sweep generates
            perm = idx + [i for i in range(M) if i not in idx]
            cal = sobol.ClosedSobol(X[:, perm], ls[:, perm], F, KiY, True)
            absum = np.zeros((L, L))
            for l in range(L):
                for j in range(L):
                    H = sobol.H_block(cal.X, cal.Phi[l, 0], cal.Phi[j, 0], 0, len(idx))
                    absum[l, j] = np.abs(cal.g0KY[l, 0]) @ (np.abs(H) @ np.abs(cal.g0KY[j, 0]))
            ref_v = cal.marginalize((0, len(idx)))['V']
            err = np.abs(v - ref_v)
            bound = 1e-10 + 1e-8 * np.abs(ref_v) + 8 * np.finfo(float).eps * absum
            assert np.all(err <= bound), f'{tag} sobol mask={m:b}: max err/bound {np.max(err / bound):.2f}, max abs err {err.max():.3e}'
        # round 2: the lattice form (lists holding at least half of a block of 2^min(6,M) subsets) against the sweep form / the one-exp-per-subset
        # kernel, which the same subsets reach through a sparse list; row-tile parts of a multi-GPU sweep
        if M >= 3:
            KL = min(6, M)
            block = int(rng.integers(0, 1 << (M - KL))) << KL
            dense = [block | lo for lo in rng.permutation(1 << KL)[: int(rng.integers((1 << KL) // 2, (1 << KL) + 1))]]
            Vd = C.sobol_contract(dX, Phi, g0KY, L, True, dense).cpu().numpy()
            Vs = np.stack([C.sobol_contract(dX, Phi, g0KY, L, True, [m]).cpu().numpy()[0] for m in dense[:4]])
            cal0 = sobol.ClosedSobol(X, ls, F, KiY, True)
            worst_sum = np.zeros((L, L))               # sum |terms| of these subsets: what the rounding of ANY evaluation order scales with
            for m in dense[:4]:
                idx = [i for i in range(M) if (m >> i) & 1]
                perm = idx + [i for i in range(M) if i not in idx]
                for l in range(L):
                    for j in range(L):
                        H = sobol.H_block(X[:, perm], cal0.Phi[l, 0][perm], cal0.Phi[j, 0][perm], 0, len(idx))
                        worst_sum[l, j] = max(worst_sum[l, j], np.abs(cal0.g0KY[l, 0]) @ (np.abs(H) @ np.abs(cal0.g0KY[j, 0])))
            bound = 1e-10 + 1e-8 * np.abs(Vs) + 16 * np.finfo(float).eps * worst_sum
            assert np.all(np.abs(Vd[:4] - Vs) <= bound), f'{tag} lattice form vs single subsets: {np.max(np.abs(Vd[:4] - Vs) / bound):.2f} of the bound'
            nparts = int(rng.integers(2, 5))
            tot = sum(C.sobol_contract(dX, Phi, g0KY, L, True, dense, None, r, nparts).cpu().numpy() for r in range(nparts))
            assert np.all(np.abs(tot - Vd) <= 1e-10 + 1e-8 * np.abs(Vd) + 16 * np.finfo(float).eps * worst_sum.max()), f'{tag} lattice parts add up'

        # round 2: errors with the MIXED rank equation against the oracle (small N only: the oracle is O(L^2 N^2) per subset and call)
        if N <= 200 and (M <= 6 or N <= 100):
            from oracle import sobol_error
            cho = gp.k_cho_mo(X, ls, F, E)
            fac = C.Factorization(C.gram(dX, None, args[0], args[1], args[2], lower_only=True, pad_to=L * N, pad_identity=True))
            refe = sobol_error.ClosedSobolWithError(X, ls, Fd, KiY, cho, is_T_partial=False)
            m0 = int(rng.integers(0, M))
            m1 = int(rng.integers(m0 + 1, M + 1))
            Ve, We, Wm = (t.cpu().numpy() for t in C.sobol_error(dX, C.dev(ls), C.dev(Fd), Phi, g0, g0KY, fac, [C.slice_mask(m0, m1)], mixed=True))
            out = refe.marginalize((m0, m1))
            # W and WMm are differences of dots of SOLVED columns (psi = K_cho^-1 ...).  With M = 1 the samples lie on a line, the integral matrices are
            # nearly rank deficient, and the float64 oracle is not self-consistent below 3e-8..3e-7 of the cancelling terms: for the full model its DIAGONAL
            # and MIXED forms (equal in exact arithmetic) differ by that much (seed 1 of random_problem(193, 1, 1): 2.6e-8).  The comparison allows for it.
            rel = max(1e-7, 4.0 * float(np.linalg.cond(cho @ cho.T)) * np.finfo(float).eps)   # psi = K_cho^-1 (...): cond(K) eps of the cancelling terms on both sides
            if m1 - m0 == M:
                full = refe.marginalize((0, M))
                floor = np.max(np.abs(full['W'] - full['WMm']) / np.maximum(full['WMm_scale'], 1e-300))
                rel = max(rel, 4.0 * float(floor))
            assert np.all(np.abs(We[0] - out['W']) <= rel * out['W_scale'] + 1e-10), \
                f"{tag} W[{m0}:{m1}]: worst err/scale {np.max(np.abs(We[0] - out['W']) / out['W_scale']):.2e} (allowed {rel:.1e}), max abs err {np.abs(We[0] - out['W']).max():.2e}"
            assert np.all(np.abs(Wm[0] - out['WMm']) <= rel * out['WMm_scale'] + 1e-10), \
                f"{tag} WMm[{m0}:{m1}]: got {Wm[0].ravel()[:4]} ref {out['WMm'].ravel()[:4]} scale {np.ravel(out['WMm_scale'])[:4]} rel {rel:.1e}"
    # round 2: independent problems in one batched call (rc_lml_grad_multi) - this problem beside two smaller random ones
    if L * N <= 600:
        others = [random_problem(max(2, int(rng.integers(2, N + 1))), M, L, seed=int(rng.integers(1 << 30)), full_F=full_F) for _ in range(2)]
        probs = [(X, Y, ls, F, E)] + others
        mp = C.LmlGradMultiPlan([C.dev(p[0]) for p in probs], [C.dev(p[1]) for p in probs], L, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
        outm = mp.unpack(mp(C.dev(np.concatenate([p[2] for p in probs])), C.dev(np.stack([p[3] for p in probs])), C.dev(np.stack([p[4] for p in probs]))).cpu().numpy())
        assert mp.info.cpu().tolist() == [0, 0, 0], tag
        for z, pz in enumerate(probs):
            rz = ref if z == 0 else gp.lml_grad_mo(*pz)
            assert_close(outm[z]['lml'], rz['lml'], what=f'{tag} multi[{z}] lml')
            for k in ('dF', 'dE', 'dls'):
                assert_close(outm[z][k], rz[k], atol=1e-10 * L * pz[0].shape[0], what=f'{tag} multi[{z}] {k}')
    cases += 1
print(f'fuzz ok: {cases} random cases in {time.time() - t0:.0f} s')
