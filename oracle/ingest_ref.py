"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's text loaders and id numbering:
  load_pairs          ncl.py:542-543 (`[[*line.strip().split()[:2], 1.0] for line in open(path) if line.strip()]`)
  number_ids          ncl.py:55-62 (enumerate(sorted(set(ids))): STRING order) / selfcf.py:281-288 (first appearance)
  pack_key            the 8-byte big-endian key the GPU path sorts instead of the strings (same order for ASCII ids)
Pinned by tests/golden/ingest.npz, which the reference's own load_data + Interaction classes produced.
"""
from __future__ import annotations

from typing import Dict, List, Tuple


def load_pairs(data: bytes) -> List[Tuple[str, str]]:
    out = []
    for line in data.decode("ascii").split("\n"):
        if line.strip():
            tok = line.strip().split()
            out.append((tok[0], tok[1]))
    return out


def number_ids(ids: List[str], order: str) -> Dict[str, int]:
    if order == "sorted":
        return {s: k for k, s in enumerate(sorted(set(ids)))}
    table: Dict[str, int] = {}
    for s in ids:
        if s not in table:
            table[s] = len(table)
    return table


def pack_key(s: str) -> int:
    b = s.encode("ascii")
    if len(b) > 8:
        raise ValueError("id longer than 8 bytes")
    return int.from_bytes(b.ljust(8, b"\0"), "big")


def pack_words(s: str, n_words: int) -> Tuple[int, ...]:
    """The key tuple of an id of up to 8 * n_words bytes (gcf_text_parse_pairs_words): big-endian 64-bit words of the
    zero-padded byte string; tuple order == byte-wise string order for ASCII ids."""
    b = s.encode("ascii")
    if len(b) > 8 * n_words:
        raise ValueError("id longer than the key")
    b = b.ljust(8 * n_words, b"\0")
    return tuple(int.from_bytes(b[8 * w: 8 * w + 8], "big") for w in range(n_words))
