"""CPU tests: the C-ABI library loads and exports every symbol include/cpimh.h declares, struct layouts match, compute
calls fail loudly without a GPU (no CPU fallback), and the host-side logic (R-shaped chain driver, H_bar, sharding,
all-reduce of the estimator sums over a 2-rank gloo group) behaves like the reference's R code."""
import ctypes as C
import math
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "cpimh.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cpimh_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import unbiased_pimh_b200 as up
    lib = up.lib()
    syms = declared_symbols()
    assert len(syms) >= 50
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.cpimh_version() == 1


def test_struct_layouts_match_header():
    """sizeof(cpimh_config) etc. as the C compiler sees them == the ctypes mirrors."""
    from unbiased_pimh_b200 import _lib as L
    src = '#include <stdio.h>\n#include "cpimh.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(cpimh_config), sizeof(cpimh_noise), sizeof(cpimh_chain_out));return 0;}'
    exe = "/tmp/cpimh_sizeof_%d" % os.getpid()
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, check=True).stdout.split()
    os.remove(exe)
    assert [int(v) for v in out] == [C.sizeof(L.Config), C.sizeof(L.Noise), C.sizeof(L.ChainOut)]


def test_oracle_and_library_share_enums(oracle):
    from unbiased_pimh_b200 import _lib as L
    assert (L.MODEL_GSS, L.MODEL_AR, L.MODEL_LEVY_SV, L.MODEL_SK, L.MODEL_LORENZ) == (oracle.MODEL_GSS, oracle.MODEL_AR, oracle.MODEL_LEVY_SV, oracle.MODEL_SK, oracle.MODEL_LORENZ)
    assert (L.INTEGRATOR_RK4, L.INTEGRATOR_DOPRI5) == (oracle.INTEGRATOR_RK4, oracle.INTEGRATOR_DOPRI5)


def test_no_cpu_fallback_and_argument_checks():
    import unbiased_pimh_b200 as up
    import torch
    y = np.zeros((5, 1))
    if not torch.cuda.is_available():
        with pytest.raises(up.CpimhError) as ei:
            up.cpf_batch(64, up.get_gss(), [], y, 2, seed=1)
        assert ei.value.code == 2                      # CPIMH_ERR_CUDA: surfaced, not swallowed
        with pytest.raises(up.CpimhError):
            up.multinomial_resampling_n_(np.full(8, 0.125), 8, e_sorted=np.linspace(0.1, 0.9, 8))
    with pytest.raises(up.CpimhError) as ei:
        up.cpf_batch(0, up.get_gss(), [], y, 2, seed=1)
    assert ei.value.code == 1 and "nparticles" in str(ei.value)
    with pytest.raises(up.CpimhError) as ei:
        up.cpf_batch(16, up.get_ar(17), [0.5, 1.0], np.zeros((5, 17)), 1, seed=1)   # get_ar(d) is compiled for every d <= 16
    assert "dimension" in str(ei.value)
    with pytest.raises(up.CpimhError) as ei:          # conditional filters: argument checks come before any CUDA call
        up.CPF(16, up.get_gss(), [], y, ref_trajectory=np.zeros((1, 6)), shuffle=1)
    assert ei.value.code == 1 and "shuffle" in str(ei.value)
    with pytest.raises(up.CpimhError) as ei:
        up.coupled_ccpf_chains_batch(up.get_gss(), [], y, 16, 4, K=3, k=5)
    assert ei.value.code == 1


def test_product_does_not_import_the_oracle():
    """The product package must never route through oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "unbiased_pimh_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "oracle.h" not in txt and "orc_" not in txt, os.path.join(dirpath, f)
    hdr = open(os.path.join(ROOT, "include", "cpimh.h")).read()
    assert "orc_" not in hdr


# ---------------------------------------------------------------- host logic
def fake_kernels(logzs, trajs, us):
    """PIMH kernels (R/pf_kernels.R:88-145) over a scripted sequence of proposals and uniforms."""
    it = {"p": 0, "u": 0}

    def prop():
        i = it["p"]; it["p"] += 1
        return trajs[i], logzs[i]

    def unif():
        i = it["u"]; it["u"] += 1
        return us[i]

    def single_kernel(cs, lp, iter, N):
        x, lz = prop(); u = unif()
        if math.log(u) < lz - lp:
            cs, lp = x, lz
        return {"chain_state": cs, "log_pdf_state": lp}

    def coupled_kernel(c1, c2, l1, l2, iter, N):
        x, lz = prop(); u = unif()
        if math.log(u) < lz - l1:
            c1, l1 = x, lz
        if math.log(u) < lz - l2:
            c2, l2 = x, lz
        return {"chain_state1": c1, "chain_state2": c2, "log_pdf_state1": l1, "log_pdf_state2": l2}

    def pimh_init(N):
        x1, l1 = prop(); x2, l2 = prop()
        return {"chain_state1": x1, "chain_state2": x2, "log_pdf_state1": l1, "log_pdf_state2": l2}
    return single_kernel, coupled_kernel, pimh_init


def test_coupled_pimh_chains_and_hbar_host_logic():
    import unbiased_pimh_b200 as up
    rng = np.random.default_rng(0)
    P = 6
    trajs = [rng.normal(size=(2, P)) for _ in range(12)]
    #        init1 init2(discarded) it1   it2   it3   it4 ...
    logzs = [-5.0, -1.0, -6.0, -7.0, -4.0, -4.5, -9.0, -3.0, -3.5, -2.0, -1.0, -0.5]
    us = [0.5, 0.9, 0.2, 0.99, 0.3, 0.6, 0.4, 0.5, 0.5, 0.5]
    sk, ck, pi = fake_kernels(logzs, trajs, us)
    c = up.coupled_pimh_chains(sk, ck, pi, 10, K=5)
    # iteration 1: log(.5)=-.69 < -6+5=-1? no -> chain1 stays; chain2 (-Inf) accepts proposal #2
    assert np.array_equal(c["samples2"][0], trajs[2][0]) and np.array_equal(c["samples1"][1], trajs[0][0])
    # iteration 3: proposal logz -4: log(.2)=-1.6 < -4+5=1 -> chain1 accepts; chain2: -1.6 < -4+6=2 accepts -> meeting at 3
    assert c["meetingtime"] == 3 and c["iteration"] == 5 and c["finished"]
    assert c["samples1"].shape == (6, P) and c["samples2"].shape == (5, P) and c["log_pdf1"].shape == (6, 1)
    for k, K in ((0, 1), (1, 3), (2, 5), (0, 5)):
        hb = up.H_bar(c, k=k, K=K)
        ref = sum(c["samples1"][t] for t in range(k, K + 1))
        for t in range(k, min(c["iteration"] - 1, c["meetingtime"] - 1) + 1):
            if c["meetingtime"] > k + 1:
                ref = ref + min(t - k + 1, K - k + 1) * (c["samples1"][t + 1] - c["samples2"][t])
        assert np.allclose(hb, ref / (K - k + 1), rtol=1e-14)
    assert up.H_bar(c, k=9, K=10) is None                 # k beyond the horizon (R prints an error and returns NULL)
    for k, K in ((0, 1), (1, 3), (0, 5)):                  # H_bar = MCMC average + BC (R/debiasing_algorithm.R:330-363)
        mcmc = sum(c["samples1"][t] for t in range(k, K + 1)) / (K - k + 1)
        assert np.allclose(up.H_bar(c, k=k, K=K), mcmc + up.BC(c, k=k, K=K), rtol=1e-14)
    assert np.allclose(up.BC(c, k=0, K=math.inf), sum(min(t + 1, math.inf) * (c["samples1"][t + 1] - c["samples2"][t]) for t in range(0, 3)))


def test_python_hbar_equals_oracle_hbar(oracle, golden):
    import unbiased_pimh_b200 as up
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 40, 100, theta=[0.5, math.sqrt(10.0)])
    for rep in range(5):
        c = oracle.coupled_pimh_chains(cfg, golden["lgssm_y"], 9, rep, K=3)
        hb = up.H_bar(c, k=1, K=3)
        r = oracle.coupled_pimh_batch(cfg, golden["lgssm_y"], 9, rep, 1, k=1, K=3)
        assert np.allclose(hb, r["hbar"][0], rtol=1e-14, atol=1e-14)


def test_shard_range_partitions_replicates():
    import unbiased_pimh_b200 as up
    for R in (1, 7, 100000):
        for G in (1, 2, 4, 8):
            ranges = [up.shard_range(R, g, G) for g in range(G)]
            assert ranges[0][0] == 0 and ranges[-1][1] == R
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(G - 1))
            assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1


GLOO_SCRIPT = textwrap.dedent("""
    import os, sys, math
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    import unbiased_pimh_b200 as up
    from oracle import pyoracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = np.load(os.path.join({root!r}, "tests", "golden", "reference_fixtures.npz"))
    cfg = O.make_config(O.MODEL_AR, 1, 1, 32, 100, theta=[0.5, math.sqrt(10.0)])
    R = 37
    lo, hi = up.shard_range(R, rank, world)
    r = O.coupled_pimh_batch(cfg, g["lgssm_y"], 123, lo, hi - lo, k=0, K=1, nthreads=2)   # the oracle stands in for the GPU shard
    P = 101
    sums = torch.zeros(2 * P + 2, dtype=torch.float64)
    sums[:P] = torch.from_numpy(r["hbar"].sum(axis=0)); sums[P:2 * P] = torch.from_numpy((r["hbar"] ** 2).sum(axis=0))
    sums[2 * P] = r["nfilters"]; sums[2 * P + 1] = hi - lo
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)           # the ONE collective of the multi-GPU path
    res = up.combine_sums(sums.numpy(), 1, 100)
    if rank == 0:
        full = O.coupled_pimh_batch(cfg, g["lgssm_y"], 123, 0, R, k=0, K=1, nthreads=2)
        assert res["nreplicates"] == R and res["nfilters"] == full["nfilters"]
        assert np.allclose(res["mean"][0], full["hbar"].mean(axis=0), rtol=1e-12, atol=1e-12)
        assert np.allclose(res["se"][0], full["hbar"].std(axis=0, ddof=1) / math.sqrt(R), rtol=1e-9, atol=1e-12)
        print("GLOO_OK")
    dist.destroy_process_group()
""")


def test_sharded_estimator_allreduce_gloo_world2(tmp_path, oracle):
    """world_size-2 gloo run of the multi-GPU host logic: contiguous replicate shards keyed by global replicate id give
    the same per-replicate results as one rank, and one all-reduce of the sums reproduces mean / s.e."""
    script = tmp_path / "gloo_job.py"
    script.write_text(GLOO_SCRIPT.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_rdata_reader_roundtrip(golden):
    """The committed fixtures came through the pure-Python RDX2 reader; spot-check shapes/values it produced."""
    assert golden["lgssm_y"].shape == (100, 1) and golden["sk_y"].shape == (100, 2) and golden["sv_y"].shape == (500, 1)
    assert abs(float(golden["lgssm_kalman_loglik"]) - (-279.5435298471)) < 1e-8
    assert np.allclose(golden["sk_varlogz"], [1.0222, 0.4971, 0.3264, 0.2434, 0.1937], atol=5e-4)
    assert np.allclose(golden["sv_theta_mle"], [0.244872, -0.275641, 0.816154, 0.093333, 0.051538], atol=1e-6)
    assert golden["sk_mt_hist"].sum(axis=1).tolist() == [10000] * 5


def test_r_shim_compiles():
    """r-pkg/src/shim.c (the .Call binding of INTEGRATION.md) is syntactically valid C against the C ABI header and a
    minimal stand-in for R's headers, and registers every hot-path entry of the reference's CallEntries table."""
    shim = os.path.join(ROOT, "r-pkg", "src", "shim.c")
    subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "r-pkg", "src", "stub"), "-I", os.path.join(ROOT, "include"), shim], check=True)
    txt = open(shim).read()
    for name in ("multinomial_resampling_n_", "coupled_multinomial_resampling_n_", "systematic_resampling_n_", "indexmatching_cpp", "dmvnorm",
                 "dmvnorm_transpose", "dmvnorm_transpose_cholesky", "rmvnorm", "rmvnorm_transpose", "rmvnorm_transpose_cholesky",
                 "rinitial_SVLevy_cpp", "rtransition_SVLevy_cpp", "dobs_SVLevy_cpp", "one_step_lorenz_vector", "create_A_"):
        assert "ENTRY(_CoupledPIMH_%s," % name in txt, name


def test_rdata_writer_round_trip_and_byte_identity(tmp_path):
    """write_rdata(): c_chains lists and data sets round-trip through the RDX2 reader, and re-serialising a file written by R
    itself reproduces it byte for byte (checked when the reference checkout is present; the version word aside)."""
    import gzip
    import unbiased_pimh_b200 as up
    ch = {"samples1": np.arange(12.0).reshape(4, 3), "samples2": np.arange(9.0).reshape(3, 3), "log_pdf1": np.arange(4.0).reshape(4, 1),
          "meetingtime": np.array([math.inf]), "iteration": np.array([3], dtype=np.int32), "finished": np.array([True])}
    obs = np.random.default_rng(0).normal(size=(5, 2))
    p = str(tmp_path / "t.RData")
    up.write_rdata(p, {"c_chains": ch, "observations": obs, "labels": ["a", "b"], "nothing": None})
    r = up.read_rdata(p)
    assert np.array_equal(r["observations"], obs) and r["labels"] == ["a", "b"] and r["nothing"] is None
    for k in ch:
        assert np.array_equal(np.asarray(r["c_chains"][k], dtype=float), np.asarray(ch[k], dtype=float)), k
    assert r["c_chains"]["samples1"].shape == (4, 3) and r["c_chains"]["iteration"].dtype == np.int32
    ref = "/root/reference/inst/reproduce/stochastic_kinetic/data/data100obs.RData"
    if os.path.exists(ref):
        orig = gzip.open(ref, "rb").read()
        up.write_rdata(p, up.read_rdata(ref))
        new = gzip.open(p, "rb").read()
        assert orig[:11] == new[:11] and orig[15:] == new[15:]
