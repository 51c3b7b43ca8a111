"""Host-side mirror logic that needs no GPU: fast_select == the dict path, AlleleIndex, H6 rounding probes."""
import numpy as np

from helpers import lut_from_db, small_case
from metamlst_b200 import api
from oracle import corc


def test_fast_select_equals_dict_path():
    for seed in (71, 72, 73):
        db, tab = small_case(seed=seed, n_reads=1500, orgs=("ecoli", "saureus"), apl=9, sub_err=0.03)
        allow, locus_of, n_loci = lut_from_db(db)
        for minscore in (80, 170, 182):
            s, c, f, _ = corc.score(tab, allow, locus_of, n_loci, minscore, 5, 50)
            index = api.AlleleIndex(tab.ref_names)
            cel = api.finish_scores(index, s, c, f, 100)
            want = [(sp, [index.name_to_tid["%s_%s_%s" % (sp, g, a)] for g, a in api.select_alleles(genes)]) for sp, genes in cel.items()]
            assert api.fast_select(index, s, c, f, 100) == want


def test_fast_select_rounding_ties():
    # two alleles whose raw averages differ but round to the same 0.1 -> lowest allele number wins (metamlst.py:244)
    index = api.AlleleIndex(["o_g_7", "o_g_3", "o_g_5", "o_h_1"])
    s = np.array([27050, 27049, 26000, 10], np.int64)
    c = np.array([200, 200, 200, 1], np.uint32)
    f = np.array([5, 9, 1, 0], np.uint32)
    cel = api.finish_scores(index, s, c, f, 100)
    assert cel["o"]["g"]["7"][2] == cel["o"]["g"]["3"][2] == 135.2  # 135.25 -> 135.2 (half-even on the binary value), 135.245 -> 135.2
    assert api.fast_select(index, s, c, f, 100) == [("o", [3, 1])]  # locus h first (record 0), then g -> allele 3
    assert list(cel["o"].keys()) == ["h", "g"] and list(cel["o"]["g"].keys()) == ["5", "7", "3"]


def test_allele_index_rejects_malformed_names():
    import pytest
    with pytest.raises(ValueError):
        api.AlleleIndex(["ecoli_adk_1", "ecoli_adk"])
    idx = api.AlleleIndex(["a_x_1", "b_y_2", "a_x_2"])
    assert list(idx.locus_of) == [0, 1, 0] and list(idx.allow_mask("b,c")) == [0, 1, 0] and list(idx.allow_mask(None)) == [1, 1, 1]


def test_sample_lanes_order_errors_and_lane_reuse(monkeypatch):
    """api.SampleLanes host logic without a GPU (contexts, index upload and the library call replaced): results come back in submission order whatever
    order the lanes finish in, an exception of one sample reaches its caller only, and the lane that raised goes back to the pool."""
    import time
    from metamlst_b200 import api, native

    class FakeCtx:
        def __init__(self, device):
            self.device, self.closed = device, False

        def close(self):
            self.closed = True

    class FakeIndex:
        def __init__(self, ctx, *a, **k):
            self.ctx = ctx

    seen = []

    def fake_type_soa(sidx, soa, **kw):
        seen.append((id(sidx), soa))
        if soa == "boom":
            raise RuntimeError("Database is broken")
        time.sleep(0.05 if soa % 2 == 0 else 0.0)   # even samples finish late
        return {"sample": soa, "kw": kw}

    monkeypatch.setattr(native, "Context", FakeCtx)
    monkeypatch.setattr(api, "SampleIndex", FakeIndex)
    monkeypatch.setattr(api, "type_soa", fake_type_soa)
    lanes = api.SampleLanes(0, None, [], None, lanes=3)
    try:
        got = lanes.map(list(range(10)), minscore=80)
        assert [g["sample"] for g in got] == list(range(10)) and all(g["kw"] == {"minscore": 80} for g in got)
        assert len({s for s, _ in seen}) <= 3   # three resident indexes serve all ten samples
        futs = [lanes.submit(1), lanes.submit("boom"), lanes.submit(3)]
        assert futs[0].result()["sample"] == 1 and futs[2].result()["sample"] == 3
        try:
            futs[1].result()
            raise AssertionError("the failing sample must raise")
        except RuntimeError as e:
            assert "broken" in str(e)
        assert [g["sample"] for g in lanes.map([5, 6, 7, 8])] == [5, 6, 7, 8]   # the lane that raised is back in the pool: nothing hangs
    finally:
        ctxs = list(lanes.ctxs)
        lanes.close()
    assert all(c.closed for c in ctxs) and lanes.ctxs == []
    try:
        api.SampleLanes(0, None, [], None, lanes=0)
        raise AssertionError("lanes=0 must be refused")
    except ValueError:
        pass
