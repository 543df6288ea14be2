"""Host-side shard planning for the multi-GPU pci.ids load.

The device list shards by vendor-id range: a cut may only fall where a TOP-LEVEL line starts
(first byte neither '\\t' nor '#'), so no vendor block spans two shards and a shard needs no
carry-in.  This is partitioning logic (a few memchr calls per cut), not part of the parse.
"""


def next_top_level_start(text, pos: int) -> int:
    """Smallest offset >= pos at which a top-level line starts (len(text) if none)."""
    n = len(text)
    if pos <= 0:
        return 0
    p = pos
    # a line start is 0 or one past a newline
    if text[p - 1:p] != b"\n":
        nl = text.find(b"\n", p)
        if nl < 0:
            return n
        p = nl + 1
    while p < n:
        c = text[p:p + 1]
        if c != b"\t" and c != b"#":
            return p
        nl = text.find(b"\n", p)
        if nl < 0:
            return n
        p = nl + 1
    return n


def plan_shards(text, nranks: int):
    """[(start, end)] * nranks, contiguous, covering the text, every start at a top-level line."""
    n = len(text)
    cuts = [0]
    for r in range(1, nranks):
        cuts.append(max(cuts[-1], next_top_level_start(text, n * r // nranks)))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(nranks)]
