"""Larger-size smoke run of the conditional / adaptive filters and the chain drivers (limits, shared-memory checks)."""
import math, numpy as np, sys, time
sys.path.insert(0, '.')
import unbiased_pimh_b200 as up
G = np.load('tests/golden/reference_fixtures.npz')
th = [0.5, math.sqrt(10.0)]
ysv = np.tile(G['sv_y'], (2, 1))[:300]
thsv = G['sv_theta_mle']
t0 = time.time()
x = up.cpf_batch(4096, up.get_levy_sv(), thsv, ysv, 8, seed=1)["trajectory"]
r = up.ccpf_batch(4096, up.get_levy_sv(), thsv, ysv, x[:4], x[4:], seed=2, exp_path=True)
print("ccpf sv N=4096 T=300 ok", r["logz"][:2], time.time() - t0)
r = up.ccpf_batch(5600, up.get_ar(1), th, G['lgssm_y'], np.zeros((2, 1, 101)), np.ones((2, 1, 101)), with_as=True, seed=2)
print("ccpf ar1 N=5600 ok", r["logz"])
try:
    up.ccpf_batch(8192, up.get_ar(1), th, G['lgssm_y'], np.zeros((2, 1, 101)), seed=2)
except up.CpimhError as e:
    print("ccpf N=8192 ->", str(e)[:120])
A = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
r = up.pf_batch(G['sk_y'], {"A_obs": A, "s2_obs": 1.0}, up.get_stochastic_kinetic(), 8192, 200, ess_threshold=0.5, seed=3)
print("pf sk N=8192 B=200 ok", r["loglik"][:3], r["resampled"].mean())
r = up.pf_batch(G['lgssm_y'], th, up.get_ar(1), 1000, 20000, ess_threshold=1.0, seed=3)
ll = r["loglik"]; print("pf ar1 B=20000 lme", np.log(np.mean(np.exp(ll - ll.max()))) + ll.max())
t0 = time.time()
g = up.coupled_ccpf_chains_batch(up.get_levy_sv(), thsv, ysv[:100], 1024, 500, K=3, seed=5, max_iterations=300, rb=True)
print("ccpf chains sv N=1024 R=500: tau mean %.2f max %d met %d  %.2fs" % (g["meetingtime"][g["meetingtime"]>0].mean(), g["meetingtime"].max(), (g["meetingtime"]>0).sum(), time.time()-t0))
g = up.coupled_ccpf_chains_batch(up.get_ar(10), th, np.random.default_rng(1).normal(size=(100, 10)), 512, 60000, K=2, seed=5, with_as=True)
print("ccpf chains ar10 R=60000 ok", g["meetingtime"].mean(), g["nfilters"])
g = up.coupled_pimh_chains_batch(up.get_ar(1), th, G['lgssm_y'], 256, 200000, K=1, seed=5, all_particles=True)
print("pimh chains R=200000 all_particles ok", g["meetingtime"].mean(), np.abs(g["hbar"].mean(0) - g["hbar_rb"].mean(0)).max())
