"""Minimal reader for ASE "ULM" trajectory files (no ASE needed).

Layout (found by probing gappy/example/ASE-GAPPY/ase.traj):
bytes 0-7 ``- of Ulm``, 8-23 a 16-byte tag, then three little-endian int64
(version, nitems, offset of an int64[nitems] table of item offsets).  Each item
is an int64 byte count followed by that many bytes of JSON; keys ending in
``.`` hold ``{"ndarray": [shape, dtype, absolute_byte_offset]}``.
"""
import json
import struct

import numpy as np


def _resolve(node, blob):
    if isinstance(node, dict):
        if set(node) == {"ndarray"}:
            shape, dtype, offset = node["ndarray"]
            count = int(np.prod(shape)) if shape else 1
            arr = np.frombuffer(blob, dtype=np.dtype(dtype).newbyteorder("<"),
                                count=count, offset=offset)
            return arr.reshape(shape).astype(np.dtype(dtype))
        return {k.rstrip("."): _resolve(v, blob) for k, v in node.items()}
    return node


def read_ulm(path):
    """Return (tag, [item dict, ...]) with ndarray references resolved."""
    blob = open(path, "rb").read()
    if blob[:8] != b"- of Ulm":
        raise ValueError("not an ULM file: %r" % path)
    tag = blob[8:24].decode().rstrip()
    _version, nitems, table = struct.unpack("<3q", blob[24:48])
    offsets = np.frombuffer(blob, dtype="<i8", count=nitems, offset=table)
    items = []
    for off in offsets:
        (nbytes,) = struct.unpack("<q", blob[off:off + 8])
        items.append(_resolve(json.loads(blob[off + 8:off + 8 + nbytes]), blob))
    return tag, items
