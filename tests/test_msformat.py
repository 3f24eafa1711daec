"""ms-format reader (fscl_b200/msformat.py; semantics defined there because the reference's ms path
is broken at the surveyed commit) and the -nulldist writer.  CPU only."""
import numpy as np
import pytest

from fscl_b200 import msformat


def _blocks(rng, n_blocks=3, n_hap=12, S=40):
    out = []
    for _ in range(n_blocks):
        pos = np.sort(np.round(rng.random(S), 6))
        hap = (rng.random((n_hap, S)) < rng.random(S)).astype(np.uint8)      # all frequencies occur (D1 mirrors the spectrum)
        out.append(dict(positions=pos, haplotypes=hap))
    out.append(dict(positions=np.zeros(0), haplotypes=np.zeros((0, 0), dtype=np.uint8)))   # segsites: 0
    return out


def test_round_trip_and_site_semantics():
    rng = np.random.default_rng(4)
    blocks = _blocks(rng)
    text = msformat.write_ms(blocks, 12)
    text = text.replace("//\n", "//\n[12](1:0.1,2:0.1);\n", 1)                # a tree line is skipped
    got = msformat.parse_ms(text)
    assert len(got) == len(blocks)
    for a, b in zip(got, blocks):
        assert np.array_equal(a["haplotypes"], b["haplotypes"]) and np.allclose(a["positions"], b["positions"], atol=1e-6)
    L = 50_000
    tab = msformat.ms_to_table(got, L)
    assert tab["chr_names"] == ["ms0", "ms1", "ms2"]                           # the empty block yields no chromosome
    for k in range(3):
        sel = tab["chr"] == k
        d = blocks[k]["haplotypes"].sum(axis=0)
        keep = (d > 0) & (d < 12)
        want_pos = np.sort((blocks[k]["positions"] * L).astype(np.int64)[keep])
        assert np.array_equal(tab["pos"][sel], want_pos)
        assert sorted(tab["count"][sel]) == sorted(d[keep]) and (tab["depth"][sel] == 12).all()
        assert (tab["folded"][sel] == 0).all()
    # sub-sample of haplotypes and folding to the MINOR count
    tf = msformat.ms_to_table(got, L, folded=True, sample_first=2, sample_size=8)
    sel = tf["chr"] == 0
    d = blocks[0]["haplotypes"][2:10].sum(axis=0)
    keep = (d > 0) & (d < 8)
    assert sorted(tf["count"][sel]) == sorted(np.minimum(d, 8 - d)[keep])
    assert (tf["folded"] == 1).all() and (tf["depth"] == 8).all() and (tf["count"] <= 4).all()


def test_malformed_input_is_rejected():
    with pytest.raises(msformat.MsFormatError):
        msformat.parse_ms("//\nsegsites: 3\npositions: 0.1 0.2\n010\n")
    with pytest.raises(msformat.MsFormatError):
        msformat.parse_ms("//\nsegsites: 2\npositions: 0.1 0.2\n01\n0x\n")
    with pytest.raises(msformat.MsFormatError):
        msformat.ms_to_table(msformat.parse_ms("//\nsegsites: 1\npositions: 0.5\n1\n1\n\n"), 1000)


def test_ms_table_feeds_the_oracle_pipeline():
    """An ms-derived table goes through the same preparation as a SNP table."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(9)
    blocks = _blocks(rng, n_blocks=2, n_hap=20, S=300)[:2]
    tab = msformat.ms_to_table(msformat.parse_ms(msformat.write_ms(blocks, 20)), 200_000)
    flat = po.oracle_flatten(po.load_snp_table_text(po.table_text(tab)))
    pts, _ = po.Oracle(flat, mode=0).scan(50_000)
    assert len(pts) >= 6 and set(pts["chr"]) == {0, 1} and np.isfinite(pts["clr"]).all()


def test_nulldist_writer_layout():
    from fscl_b200 import api
    pts = np.zeros(2, dtype=api.PT_DTYPE)
    pts["chr"] = [0, 0]; pts["sweep_pos"] = [100, 200]; pts["clr"] = [3.14159, 0.5]; pts["lalpha"] = [-8.0, -2.0]
    pts["permute_p"] = [1, 2]; pts["permute_n"] = [3, 2]
    clr = np.zeros((2, api.NULL_SAVE), dtype=np.float32)
    clr[0, :3] = [2.5, 0.25, 1.0]; clr[1, :2] = [9.0, 0.0]
    text = api.format_nulldist(pts, ["chrA"], clr).splitlines()
    assert text[0].startswith("chr\tpos\tCLR\talpha\tp\tn\t0.0000\t0.0001") and text[0].count("\t") == 5 + api.NULL_SAVE
    assert text[1] == "chrA\t100\t3.142\t3.355e-04\t1\t3\t0.25\t1.00\t2.50"
    assert text[2] == "chrA\t200\t0.500\t1.353e-01\t2\t2\t0.00\t9.00"
