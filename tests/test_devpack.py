"""The torch packer used by bench.py (device-agnostic, run here on CPU) must equal the generic host packer."""
import numpy as np
import pytest

from metamlst_b200 import devpack, packing, synth


@pytest.mark.parametrize("L,maxd", [(100, 8000), (150, 8000), (70, 40), (50, None)])
def test_devpack_equals_host_packer(L, maxd):
    db = synth.make_db(("ecoli", "saureus"), alleles_per_locus=5, n_profiles=10, seed=61)
    kw = dict(read_len=L, seed=61, K=3, frac_clip=0.2, frac_indel=0.2, n_frac=0.04, org_props=(0.7, 0.3))
    core = synth.gen_core(db, 900, **kw)
    tab = synth.make_sample(db, 900, **kw).sorted_by_coord()
    want = packing.pack_table(tab, 20, maxd)
    got = devpack.pack_cores(db, [core], 20, maxd).to_host(pinned=False)
    assert want.orig_idx is None
    assert got.p_recs.tobytes() == want.p_recs.tobytes()
    for f in ("tid", "as0", "xm3", "qlen", "p_pos", "p_row_off", "p_reflen", "p_as", "p_xm", "planes", "contig_start"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    assert got.max_row_words == want.max_row_words and got.n_dropped_by_cap == want.n_dropped_by_cap


def test_chunked_generation_shares_the_strain():
    db = synth.make_db(("ecoli",), alleles_per_locus=5, n_profiles=10, seed=62)
    a = synth.gen_core(db, 100, 100, seed=1, strain_seed=9)
    b = synth.gen_core(db, 100, 100, seed=2, strain_seed=9)
    assert a["strain_seqs"][0].tobytes() == b["strain_seqs"][0].tobytes() and a["truth"]["st"] == b["truth"]["st"]
    assert not np.array_equal(a["start"].numpy(), b["start"].numpy())
