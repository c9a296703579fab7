"""Rebuild the 31-base window of a site of the real panel (data/human_sites_n10.fa.gz, tests/golden/shared/sites300.fa)
from its two records.  Test / tool infrastructure only.

A record is the N-joined list of the site's 19-mers that survived the panel's uniqueness filter: 3 to 13 of the 13
19-mers of the 31-base window around the SNP, in window order but with gaps.  Every one of them contains the SNP base
(index 15 of the window), the ref and the var record differ in that base only."""


def _merge(kmers):
    """Overlap-merge window-ordered k-mers; returns the span they cover, or None if two of them do not overlap."""
    w = kmers[0]
    for km in kmers[1:]:
        for d in range(1, len(km)):
            if w[len(w) - (len(km) - d):] == km[:len(km) - d]:
                w += km[len(km) - d:]
                break
        else:
            return None
    return w


def site_window(ref_record, var_record, rng, window=31):
    """(ref_window, alt_base) with every k-mer of ref_record inside ref_window and every k-mer of var_record inside the
    window with alt_base at its centre; None when the records do not pin the window down uniquely."""
    rk, vk = ref_record.split("N"), var_record.split("N")
    a, b = _merge(rk), _merge(vk)
    half = window // 2
    if not a or not b or len(a) > window or len(b) > window:
        return None
    found = []
    for s in range(-(window - 1), window):                   # b starts s bases after a
        lo, hi = max(0, s), min(len(a), s + len(b))
        if hi - lo < 8:
            continue
        mism = [i for i in range(lo, hi) if a[i] != b[i - s]]
        if len(mism) == 1:
            found.append((s, mism[0]))
    if len(found) != 1:
        return None
    s, pa = found[0]
    oa = half - pa                                            # where a starts in the window
    ob = oa + s
    if oa < 0 or ob < 0 or oa + len(a) > window or ob + len(b) > window:
        return None
    w = [None] * window
    for i, ch in enumerate(b):
        w[ob + i] = ch
    for i, ch in enumerate(a):
        w[oa + i] = ch
    alt = b[half - ob]
    w = "".join(ch if ch else rng.choice("ACGT") for ch in w)
    v = w[:half] + alt + w[half + 1:]
    if any(k not in w for k in rk) or any(k not in v for k in vk) or alt == w[half]:
        return None
    return w, alt


def panel_windows(lines, rng, window=31):
    """[(name, ref_window, alt_base) or (name, None, None)] for the records in `lines` (header, seq, header, seq, ...)."""
    out = []
    for i in range(0, len(lines) - 3, 4):
        name = lines[i][1:].split()[0]
        r = site_window(lines[i + 1].strip(), lines[i + 3].strip(), rng, window)
        out.append((name, r[0], r[1]) if r else (name, None, None))
    return out
