"""(f3) A self-describing container for the coded latents, with ONE stream set per channel group so that a subset of
the tasks can be decoded without touching the others' bytes.

The reference has no on-disk format: its notebook dumps the raw concatenated strings without lengths
(/root/reference/src/check_bpp.ipynb:167), `compress` returns nested Python lists
(/root/reference/src/models/multi_task_compressor.py:507-534), and the Disjoint / Shared models code all channel groups
in one y string per image, so "store the tasks you need" (the paper's use case, section III) is not realised.  Here the
y latent is coded group by group (the groups of `_rate_groups()`: one per task for -m 3, tasks + "shared" for -m 4,
a single group for -m 1 / -m 2); the hyper-latent z is one stream set shared by all groups.

Layout (little endian):
    magic "MMNC" | u8 version | u8 model kind | u16 groups | u32 images | u16 M | u16 N | u16 z_h | u16 z_w | u16 y_h | u16 y_w
    per group : u16 name length | name (utf-8) | u16 first channel | u16 channel count
    lengths   : u32 x images for z, then u32 x images for every group (in table order)
    payload   : the strings in the same order, back to back
Per-image strings are CompressAI-format rANS streams (bit-exact with RansEncoder), so a stream cut out of a container
decodes with CompressAI's own decoder given the same tables and indexes.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MAGIC = b"MMNC"
VERSION = 1
_HEAD = struct.Struct("<4sBBHIHHHHHH")


class Container:
    def __init__(self, kind: int, M: int, N: int, z_shape: Tuple[int, int], y_shape: Tuple[int, int],
                 groups: Sequence[Tuple[str, int, int]], z_strings: List[bytes], y_strings: Dict[str, List[bytes]]):
        self.kind, self.M, self.N, self.z_shape = int(kind), int(M), int(N), (int(z_shape[0]), int(z_shape[1]))
        self.y_shape = (int(y_shape[0]), int(y_shape[1]))
        self.groups = [(str(n), int(a), int(c)) for n, a, c in groups]
        self.z_strings, self.y_strings = list(z_strings), {k: list(v) for k, v in y_strings.items()}
        self.n_images = len(self.z_strings)
        for name, _, _ in self.groups:
            if len(self.y_strings.get(name, ())) != self.n_images:
                raise ValueError(f"group {name!r}: expected {self.n_images} strings")

    # ------------------------------------------------------------------ sizes
    def payload_bytes(self, groups: Optional[Sequence[str]] = None) -> int:
        """Coded bytes of z plus the named groups (all groups when None): what a reader of those tasks has to fetch."""
        names = [g[0] for g in self.groups] if groups is None else list(groups)
        return sum(map(len, self.z_strings)) + sum(len(s) for n in names for s in self.y_strings[n])

    # ------------------------------------------------------------------ bytes
    def to_bytes(self) -> bytes:
        out = [_HEAD.pack(MAGIC, VERSION, self.kind, len(self.groups), self.n_images, self.M, self.N, *self.z_shape,
                          *self.y_shape)]
        for name, first, count in self.groups:
            raw = name.encode("utf-8")
            out.append(struct.pack("<H", len(raw)) + raw + struct.pack("<HH", first, count))
        sets = [self.z_strings] + [self.y_strings[g[0]] for g in self.groups]
        for strings in sets:
            out.append(np.fromiter((len(s) for s in strings), dtype="<u4", count=self.n_images).tobytes())
        for strings in sets:
            out.extend(strings)
        return b"".join(out)

    @classmethod
    def from_bytes(cls, blob: bytes, groups: Optional[Sequence[str]] = None) -> "Container":
        """Parses a container; with `groups`, only those groups' strings are sliced out of the payload (the other
        groups are absent from the result and their bytes are never copied)."""
        mv = memoryview(blob)
        if len(mv) < _HEAD.size:
            raise ValueError("truncated container")
        magic, version, kind, n_groups, n_images, M, N, zh, zw, yh, yw = _HEAD.unpack_from(mv, 0)
        if magic != MAGIC:
            raise ValueError("not an MMNC container")
        if version != VERSION:
            raise ValueError(f"unsupported container version {version}")
        pos = _HEAD.size
        table = []
        for _ in range(n_groups):
            (ln,) = struct.unpack_from("<H", mv, pos)
            name = bytes(mv[pos + 2: pos + 2 + ln]).decode("utf-8")
            first, count = struct.unpack_from("<HH", mv, pos + 2 + ln)
            table.append((name, first, count))
            pos += 2 + ln + 4
        n_sets = 1 + n_groups
        need = pos + 4 * n_images * n_sets
        if len(mv) < need:
            raise ValueError("truncated container (length table)")
        lens = np.frombuffer(mv, dtype="<u4", count=n_images * n_sets, offset=pos).reshape(n_sets, n_images).astype(np.int64)
        pos = need
        if len(mv) != pos + int(lens.sum()):
            raise ValueError("container size does not match its length table")
        starts = pos + np.concatenate([[0], np.cumsum(lens.reshape(-1))[:-1]]).reshape(n_sets, n_images)
        wanted = None if groups is None else set(groups)
        unknown = (wanted or set()) - {t[0] for t in table}
        if unknown:
            raise KeyError(f"no such group(s) in the container: {sorted(unknown)}")

        def cut(row):
            return [bytes(mv[int(a): int(a) + int(n)]) for a, n in zip(starts[row], lens[row])]

        y = {name: cut(1 + i) for i, (name, _, _) in enumerate(table) if wanted is None or name in wanted}
        kept = [t for t in table if wanted is None or t[0] in wanted]
        return cls(kind, M, N, (zh, zw), (yh, yw), kept, cut(0), y)
