"""Read sharding for multi-GPU runs: reads are independent units, so a FASTA is cut into contiguous shards, one per
rank; there is no collective on the data path.  Outputs are concatenated in rank order (= input order)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def read_lengths(text: bytes) -> List[int]:
    """Lengths of the reads of a FASTA text (handle_one_file.c:201-269: '>' starts a read, CR/LF ignored)."""
    out, cur, seen = [], 0, False
    for line in text.split(b"\n"):
        if line.startswith(b">"):
            if seen:
                out.append(cur)
            seen, cur = True, 0
        else:
            cur += len(line.rstrip(b"\r"))
    if seen:
        out.append(cur)
    return out


def plan_shards(lengths: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) read ranges, one per rank, balanced by bases (the DP cost grows with length)."""
    n, total = len(lengths), sum(lengths)
    bounds, acc, r = [0], 0, 1
    for i, L in enumerate(lengths):
        acc += L
        while r < world and acc >= total * r / world:
            bounds.append(i + 1)
            r += 1
    while len(bounds) < world:
        bounds.append(n)
    bounds.append(n)
    return [(bounds[k], max(bounds[k], bounds[k + 1])) for k in range(world)]


def merge_outputs(parts: Sequence[bytes]) -> bytes:
    """Rank-ordered concatenation: shards are contiguous and ordered, so this is the input order."""
    return b"".join(parts)
