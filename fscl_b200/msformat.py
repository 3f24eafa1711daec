"""Hudson `ms` output -> SNP tables for the scan (host side).

The reference's ms path (`ms-input.c` + a bison/flex parser) does not work at the
surveyed commit: `msparser_block()` is called without its scanner handle
(ms-input.c:42,108 vs ms-parser.y:258), chromosome limits are never filled in, and
folded minor-allele sites are stored with obs_freq = 0 (ms-input.c:74-77).  This
module implements the INTENDED semantics, stated here because nothing pins them:

* a block starts at a line beginning with `//`; `segsites: N`; optional `prob: x`;
  `positions: f1 ... fN` (fractions of the segment); then one [01]+ row per haplotype;
  a blank line ends the block (ms-scanner.lex:11-34, ms-parser.y:96-246).  Newick tree
  lines between `//` and `segsites:` are skipped.
* derived count of a site = number of '1' among haplotypes [first, first+size)
  (`--ms-sample-first`, `--ms-sample-size`; size 0 = all remaining); sites with count
  0 or size are dropped (ms-input.c:69-72, 127-131);
* position = int(fraction * segment_length) (ms-input.c:64, 123; C truncation);
* `--ms-folded`: the site is folded and carries the MINOR count (the fix of the defect above);
* every block becomes its own chromosome `ms<k>` of ONE table, so all blocks are scanned in
  one launch against a background spectrum estimated from all blocks together
  (the reference's two-pass design, ms-input.c:25-91 then :93-151).
"""
from __future__ import annotations

import numpy as np


class MsFormatError(ValueError):
    pass


def parse_ms(text):
    """-> list of blocks: dict(positions=float64[S], haplotypes=uint8[n_hap][S])"""
    blocks, cur = [], None
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        line = lines[i].strip()
        if line.startswith("//"):
            cur = dict(segsites=None, positions=None, rows=[])
            blocks.append(cur)
        elif cur is not None:
            if line.startswith("segsites:"):
                cur["segsites"] = int(line.split(":", 1)[1])
            elif line.startswith("positions:"):
                cur["positions"] = np.array(line.split(":", 1)[1].split(), dtype=np.float64)
            elif line.startswith("prob:") or line.startswith("[") or line.startswith("("):
                pass
            elif line and set(line) <= {"0", "1"}:
                cur["rows"].append(line)
            elif line == "":
                if cur["rows"] or cur["segsites"] == 0:
                    cur = None
            else:
                raise MsFormatError(f"line {i + 1}: unexpected text in an ms block: {line[:40]!r}")
        i += 1
    out = []
    for k, b in enumerate(blocks):
        S = b["segsites"]
        if S is None:
            raise MsFormatError(f"block {k}: no segsites line")
        if S == 0:
            out.append(dict(positions=np.zeros(0), haplotypes=np.zeros((0, 0), dtype=np.uint8)))
            continue
        if b["positions"] is None or len(b["positions"]) != S:
            raise MsFormatError(f"block {k}: positions do not match segsites")
        if not b["rows"] or any(len(r) != S for r in b["rows"]):
            raise MsFormatError(f"block {k}: haplotype rows do not match segsites")
        hap = np.frombuffer("".join(b["rows"]).encode(), dtype=np.uint8).reshape(len(b["rows"]), S) - ord("0")
        out.append(dict(positions=b["positions"], haplotypes=hap))
    return out


def ms_to_table(blocks, segment_length, folded=False, sample_first=0, sample_size=0):
    """One table, one chromosome per block (see module docstring)."""
    if segment_length <= 0:
        raise MsFormatError("--ms-segment-length must be positive")
    names, chr_, pos, cnt, dep, fol = [], [], [], [], [], []
    for k, b in enumerate(blocks):
        hap = b["haplotypes"]
        if hap.size == 0:
            continue
        size = sample_size if sample_size > 0 else hap.shape[0] - sample_first
        if size <= 0 or sample_first + size > hap.shape[0]:
            raise MsFormatError(f"block {k}: sample range exceeds {hap.shape[0]} haplotypes")
        d = hap[sample_first:sample_first + size].sum(axis=0).astype(np.int64)
        keep = (d > 0) & (d < size)
        p = (b["positions"] * segment_length).astype(np.int64)[keep]          # C truncation of a positive double
        d = d[keep]
        if folded:
            d = np.minimum(d, size - d)
        order = np.argsort(p, kind="stable")
        if len(p):
            names.append(f"ms{k}")
            chr_.append(np.full(len(p), len(names) - 1))
            pos.append(p[order]); cnt.append(d[order]); dep.append(np.full(len(p), size))
            fol.append(np.full(len(p), 1 if folded else 0))
    if not names:
        raise MsFormatError("no usable sites in the ms input")
    cat = lambda a: np.concatenate(a).astype(np.int32)
    return dict(chr_names=names, chr=cat(chr_), pos=cat(pos), count=cat(cnt), depth=cat(dep), folded=cat(fol))


def write_ms(blocks, n_hap, header="ms synthetic"):
    """Text of an ms run holding `blocks` (for tests and synthetic inputs)."""
    out = [f"{header}", "1 2 3", ""]
    for b in blocks:
        out.append("//")
        S = len(b["positions"])
        out.append(f"segsites: {S}")
        if S:
            out.append("positions: " + " ".join(f"{x:.6f}" for x in b["positions"]))
            out.extend("".join(map(str, row)) for row in b["haplotypes"])
        out.append("")
    return "\n".join(out) + "\n"
