"""Test helper: writes a small HDF5 file in the layout libhdf5 / h5py produce by default (superblock version 0,
old-style root group = symbol table + version-1 B-tree + local heap + one SNOD, version-1 object headers,
contiguous or chunked+deflate datasets, version-1 attribute messages).  Only what the ``.athdf`` reader of
``mahakala_b200.grmhd._hdf5_min`` has to understand; written independently of that reader's code paths (it shares
no helper with it) so that a reader/writer pair of bugs cannot cancel silently on the structures checked against
the MATLAB-written file in tests/test_host_cpu.py."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _datatype(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        exp_bits, mant_bits, bias = (11, 52, 1023) if dt.itemsize == 8 else (8, 23, 127)
        bits = 0x20 | ((8 * dt.itemsize - 1) << 8)            # implied mantissa msb, sign bit position
        head = struct.pack("<B3sI", 0x11, bits.to_bytes(3, "little"), dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, mant_bits, exp_bits, 0, mant_bits, bias)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0
        return struct.pack("<B3sI", 0x10, bits.to_bytes(3, "little"), dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<B3sI", 0x13, (1).to_bytes(3, "little"), dt.itemsize)      # null-padded ASCII
    raise TypeError(dt)


def _dataspace(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _message(mtype, body):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _attribute(name, value):
    value = np.asarray(value)
    nm = name.encode() + b"\x00"
    dt, ds = _datatype(value.dtype), _dataspace(value.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds)
    return _message(0x0C, body + np.ascontiguousarray(value).tobytes())


def write_hdf5(path, datasets, attrs=None, chunked=(), userblock=0):
    """datasets: {name: ndarray}; attrs: {name: ndarray} on the root group; names in ``chunked`` are stored as
    deflate-compressed chunks (chunk = half the extent along axis 0); ``userblock`` (0 or a power of two >= 512)
    puts the superblock behind a user block, as MATLAB does."""
    attrs = attrs or {}
    names = sorted(datasets)
    out = bytearray(b"\x00" * 96)                              # superblock placeholder

    def append(b, align=8):
        out.extend(b"\x00" * (-len(out) % align))
        pos = len(out)
        out.extend(b)
        return pos

    # local heap data segment: offset 0 holds the empty name
    heap = bytearray(b"\x00" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        heap.extend(_pad8(n.encode() + b"\x00"))
    headers = {}
    for n in names:
        a = np.ascontiguousarray(datasets[n])
        msgs = [_message(0x01, _dataspace(a.shape)), _message(0x03, _datatype(a.dtype))]
        if n in chunked:
            c0 = max(1, a.shape[0] // 2)
            cshape = (c0,) + a.shape[1:]
            entries = []
            for o in range(0, a.shape[0], c0):
                chunk = np.zeros(cshape, dtype=a.dtype)
                part = a[o:o + c0]
                chunk[:part.shape[0]] = part
                raw = zlib.compress(chunk.tobytes(), 4)
                entries.append((o, append(raw), len(raw)))
            node = struct.pack("<4sBBHQQ", b"TREE", 1, 0, len(entries), UNDEF, UNDEF)
            for o, addr, nbytes in entries:
                node += struct.pack("<II", nbytes, 0) + struct.pack("<Q", o) + b"".join(struct.pack("<Q", 0) for _ in a.shape[1:]) \
                    + struct.pack("<Q", 0) + struct.pack("<Q", addr)
            node += struct.pack("<II", 0, 0) + struct.pack("<Q", a.shape[0]) + b"".join(struct.pack("<Q", 0) for _ in a.shape[1:]) \
                + struct.pack("<Q", 0)
            btree = append(node)
            layout = struct.pack("<BBB", 3, 2, a.ndim + 1) + struct.pack("<Q", btree) + \
                b"".join(struct.pack("<I", c) for c in cshape) + struct.pack("<I", a.dtype.itemsize)
            msgs.append(_message(0x0B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<II", 4, 0)))
        else:
            data = append(a.tobytes())
            layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", data, a.nbytes)
        msgs.append(_message(0x08, layout))
        headers[n] = append(_object_header(msgs))
    heap_data = append(bytes(heap))
    heap_hdr = append(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), UNDEF, heap_data))
    snod = struct.pack("<4sBBH", b"SNOD", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], headers[n], 0, 0)
    snod_pos = append(snod)
    tree = struct.pack("<4sBBHQQ", b"TREE", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_pos, name_off[names[-1]])
    tree_pos = append(tree)
    root_msgs = [_message(0x11, struct.pack("<QQ", tree_pos, heap_hdr))] + [_attribute(k, v) for k, v in attrs.items()]
    root = append(_object_header(root_msgs))
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 16, 16, 0)
    sb += struct.pack("<QQQQ", userblock, UNDEF, len(out), UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_pos, heap_hdr)
    out[0:len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(b"\x00" * userblock)
        fh.write(bytes(out))


def write_hdf5_v2(path, datasets, attrs=None):
    """The same content in the 'latest' file-format flavour (h5py libver='latest'): superblock version 2, version-2
    object headers ("OHDR", 2-byte chunk size, no times), a new-style root group with compact link messages,
    version-2 dataspaces, version-3 attribute messages, contiguous version-3 layouts.  Checksums are written as
    zero (the reader does not verify them)."""
    attrs = attrs or {}
    out = bytearray(b"\x00" * 48)                              # superblock: 12 + 4 * 8 + 4 bytes

    def append(b, align=8):
        out.extend(b"\x00" * (-len(out) % align))
        pos = len(out)
        out.extend(b)
        return pos

    def dataspace2(shape):
        return struct.pack("<BBBB", 2, len(shape), 0, 1 if len(shape) else 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)

    def msg2(mtype, body):
        return struct.pack("<BHB", mtype, len(body), 0) + body

    def ohdr2(messages):
        body = b"".join(messages)
        return b"OHDR" + struct.pack("<BB", 2, 0x01) + struct.pack("<H", len(body)) + body + b"\x00" * 4

    def attribute3(name, value):
        value = np.asarray(value)
        nm = name.encode() + b"\x00"
        dt, ds = _datatype(value.dtype), dataspace2(value.shape)
        body = struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(ds), 0) + nm + dt + ds
        return msg2(0x0C, body + np.ascontiguousarray(value).tobytes())

    links = []
    for n in sorted(datasets):
        a = np.ascontiguousarray(datasets[n])
        data = append(a.tobytes())
        layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", data, a.nbytes)
        hdr = append(ohdr2([msg2(0x01, dataspace2(a.shape)), msg2(0x03, _datatype(a.dtype)), msg2(0x08, layout)]))
        nm = n.encode()
        links.append(msg2(0x06, struct.pack("<BBB", 1, 0, len(nm)) + nm + struct.pack("<Q", hdr)))
    # link info message: version 0, flags 0, fractal heap address and name-index B-tree address undefined (compact)
    linfo = msg2(0x02, struct.pack("<BB", 0, 0) + struct.pack("<QQ", UNDEF, UNDEF))
    root = append(ohdr2([linfo] + links + [attribute3(k, v) for k, v in attrs.items()]))
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, UNDEF, len(out), root) + b"\x00" * 4
    out[0:len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(out))
