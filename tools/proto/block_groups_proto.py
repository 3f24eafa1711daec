"""Numpy statement of permute_kernel's grouping rule (fscl_kernels.cu): consecutive blocks of a permutation's list that touch
no common site are applied as one set of element swaps; a block that overlaps itself goes alone through the closed form of
the reference's chained swap.  `apply_sequential` is snp_block_permute's swap loop (scan-chromosome.c:365-384) on a class
array; `apply_grouped` is what the kernel does.  tests/test_host_logic.py compares the two on random, heavily overlapping lists."""
import numpy as np


def apply_sequential(a, blocks):
    a = a.copy()
    for i, j, L in blocks:
        if i == j or L <= 0:
            continue
        for t in range(L):                      # element by element, as the reference: overlapping ranges chain
            a[i + t], a[j + t] = a[j + t], a[i + t]
    return a


def _meet(x, lx, y, ly):
    return x < y + ly and y < x + lx


def group_sizes(blocks):
    """How many blocks each barrier-to-barrier step of the kernel takes (warp 0's rule over the next 32 blocks)."""
    out, b = [], 0
    while b < len(blocks):
        lanes = blocks[b:b + 32]
        triv = [i == j or L <= 0 for i, j, L in lanes]
        selfs = [(not t) and abs(i - j) < L for (i, j, L), t in zip(lanes, triv)]
        cnt = len(lanes)
        for l in range(1, len(lanes)):
            i, j, L = lanes[l]
            hit = selfs[l] or selfs[0]
            for m in range(min(l, 31)):
                im, jm, lm = lanes[m]
                if not triv[m] and not triv[l]:
                    hit = hit or _meet(i, L, im, lm) or _meet(i, L, jm, lm) or _meet(j, L, im, lm) or _meet(j, L, jm, lm)
            if hit:
                cnt = l
                break
        out.append(cnt)
        b += cnt
    return out


def apply_grouped(a, blocks):
    a = a.copy()
    b = 0
    for cnt in group_sizes(blocks):
        grp = blocks[b:b + cnt]
        i, j, L = grp[0]
        if i != j and L > 0 and abs(i - j) < L:             # alone by the rule above: closed form through a scratch copy
            lo, d = min(i, j), abs(i - j)
            sc = a[lo:lo + L + d].copy()
            a[lo:lo + L] = sc[d:d + L]
            for r in range(d):
                a[lo + r + d * ((L - r + d - 1) // d)] = sc[r]
        else:
            snap = a.copy()                                  # every swap of the group reads the state before the group
            for i, j, L in grp:
                if i == j or L <= 0:
                    continue
                a[i:i + L] = snap[j:j + L]
                a[j:j + L] = snap[i:i + L]
        b += cnt
    return a
