"""Small ragged-size pass over every kernel family (bootstrap / conditional / adaptive filters, chains): a quick crash and
bounds smoke run; every result here is also covered by the parity tests."""
import math, numpy as np, sys
sys.path.insert(0, '.')
import unbiased_pimh_b200 as up
G = np.load('tests/golden/reference_fixtures.npz')
y = G['lgssm_y'][:30]
th = [0.5, math.sqrt(10.0)]
for N in (5, 37, 100, 257, 1000):
    for sp in (2, 3):
        r = up.cpf_batch(N, up.get_ar(1), th, y, 5, seed=1, store_paths=sp)
    r = up.cpf_batch(N, up.get_ar(1), th, y, 3, seed=1, shuffle=1)
    x = r['trajectory']
    if N >= 5:
        up.ccpf_batch(N, up.get_ar(1), th, y, x[:1], x[1:2], with_as=True, seed=2, exp_path=True, ancestors=True)
        up.ccpf_batch(N, up.get_ar(1), th, y, x[:2], with_as=False, seed=2)
        up.pf_batch(y, th, up.get_ar(1), N, 3, ess_threshold=0.5, systematic=True, seed=3)
        up.pf_batch(y, th, up.get_ar(1), N, 3, ess_threshold=1.0, seed=3, ancestors=True)
ysv = G['sv_y'][:20]
up.cpf_batch(300, up.get_levy_sv(), G['sv_theta_mle'], ysv, 4, seed=1)
pb = up.PfBatch(up.get_levy_sv(), G['sv_theta_mle'], ysv, 300, 4); pb.all_particles(True); pb.run(4, 3); pb.fetch_exp_paths(); pb.close()
up.coupled_pimh_chains_batch(up.get_ar(1), th, y, 50, 20, K=2, seed=4, all_particles=True)
up.coupled_ccpf_chains_batch(up.get_ar(1), th, y, 50, 20, K=2, seed=4, with_as=True, rb=True)
A = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
up.cpf_batch(100, up.get_stochastic_kinetic(), {"A_obs": A, "s2_obs": 1.0}, G['sk_y'][:10], 3, seed=1)
print("sanity driver done")
