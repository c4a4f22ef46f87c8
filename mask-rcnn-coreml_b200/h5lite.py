"""Minimal HDF5 reader (and a writer for the same subset) for Keras weight files -- SURVEY.md section 8 f4.

The reference converts Matterport's `mask_rcnn_coco.h5` (Keras `save_weights`, written through h5py with its default
"earliest" library bounds) into the three .mlmodel artefacts (Conversion/task.py:163, DownloadCommand.swift:15,32).
Neither h5py nor libhdf5 is in this image, so this module reads the file format directly, from the published HDF5 File
Format Specification (version 1.1/2.0 structures that such files use):

  superblock v0/v1 -> root symbol-table entry -> object header v1 (+ continuation blocks) ->
  groups: symbol-table message -> v1 B-tree (node type 0) + local heap + symbol-table nodes (SNOD)
  datasets: dataspace v1/v2, datatype (IEEE float 16/32/64, fixed-point 8..64), layout v1..v3
            (compact, contiguous, chunked through a v1 B-tree of node type 1; deflate + shuffle filters)

Not supported (raises H5Error naming the feature): superblock v2/v3 with v2 object headers / link messages ("latest"
library bounds), variable-length and compound types, external storage, other filters.  Attributes are skipped: a Keras
weight file is addressed by its dataset paths (`<layer>/<layer>/kernel:0`), `layer_names` / `weight_names` only repeat
them.

NOT VALIDATED AGAINST A REAL FILE: no HDF5 file or library is reachable offline.  The tests build files with `write()`
below -- an independent encoder of the same structures, including multi-level group B-trees, continuation blocks and
chunked + deflated datasets -- so they check self-consistency against the specification as understood here, not
interoperability; DESIGN.md says so.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# =====================================================================================================================
# reader
# =====================================================================================================================
class File:
    """read-only view: `File(bytes)`; `.datasets()` -> {path: numpy array}; `.groups()`; `[path]`."""

    def __init__(self, data):
        self.buf = memoryview(data)
        start = self._find_superblock()
        self.base = 0
        self._superblock(start)

    # ---- low level --------------------------------------------------------------------------------------------------
    def _need(self, pos, n, what):
        if pos < 0 or pos + n > len(self.buf):
            raise H5Error(f"truncated file: {what} at {pos}+{n} beyond {len(self.buf)} bytes")

    def _u(self, pos, n, what="field"):
        self._need(pos, n, what)
        return int.from_bytes(self.buf[pos:pos + n], "little")

    def _off(self, pos):
        return self._u(pos, self.so, "offset")

    def _len(self, pos):
        return self._u(pos, self.sl, "length")

    def _find_superblock(self):
        pos = 0
        while pos + 8 <= len(self.buf):             # the superblock may sit at 0, 512, 1024, ... (user block in front)
            if bytes(self.buf[pos:pos + 8]) == SIGNATURE:
                return pos
            pos = 512 if pos == 0 else pos * 2
        raise H5Error("not an HDF5 file (signature not found)")

    def _superblock(self, p):
        version = self._u(p + 8, 1)
        if version not in (0, 1):
            raise H5Error(f"superblock version {version} (v2 object headers / link messages) is not supported; "
                          "re-save the weights with h5py's default libver='earliest'")
        self.so, self.sl = self._u(p + 13, 1), self._u(p + 14, 1)
        if self.so not in (4, 8) or self.sl not in (4, 8):
            raise H5Error(f"unsupported offset/length sizes {self.so}/{self.sl}")
        self.leaf_k, self.internal_k = self._u(p + 16, 2), self._u(p + 18, 2)
        q = p + 24 + (4 if version == 1 else 0)       # v1 adds indexed-storage K (2) + reserved (2)
        self.base = self._off(q)
        q += 4 * self.so                              # base, free-space info, end of file, driver info
        self.root = self._symbol_entry(q)

    def _addr(self, a):
        return a + self.base

    def _symbol_entry(self, p):
        """link name offset, object header address, cache type, reserved, 16 bytes scratch"""
        name_off, header = self._off(p), self._off(p + self.so)
        cache = self._u(p + 2 * self.so, 4)
        scratch = p + 2 * self.so + 8
        e = {"name_off": name_off, "header": header, "cache": cache}
        if cache == 1:
            e["btree"], e["heap"] = self._off(scratch), self._off(scratch + self.so)
        return e

    @property
    def entry_size(self):
        return 2 * self.so + 4 + 4 + 16

    # ---- object headers ---------------------------------------------------------------------------------------------
    def _messages(self, addr):
        """[(type, flags, payload memoryview)] of a version-1 object header incl. continuation blocks."""
        p = self._addr(addr)
        self._need(p, 16, "object header")
        if bytes(self.buf[p:p + 4]) == b"OHDR":
            raise H5Error("version-2 object header is not supported (file written with libver='latest')")
        version = self._u(p, 1)
        if version != 1:
            raise H5Error(f"object header version {version} at {addr}")
        count, size = self._u(p + 2, 2), self._u(p + 8, 4)
        blocks = [(p + 16, size)]                      # 12 bytes of prefix padded to 16
        out = []
        while blocks and len(out) < count:
            q, n = blocks.pop(0)
            end = q + n
            self._need(q, n, "object header block")
            while q + 8 <= end and len(out) < count:
                mtype, msize, flags = self._u(q, 2), self._u(q + 2, 2), self._u(q + 4, 1)
                body = self.buf[q + 8:q + 8 + msize]
                if len(body) != msize:
                    raise H5Error("truncated header message")
                if mtype == 0x0010:                    # continuation: offset, length
                    blocks.append((self._addr(int.from_bytes(body[:self.so], "little")),
                                   int.from_bytes(body[self.so:self.so + self.sl], "little")))
                out.append((mtype, flags, body))
                q += 8 + msize
        return out

    # ---- groups -----------------------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr, offset):
        p = self._addr(heap_addr)
        if bytes(self.buf[p:p + 4]) != b"HEAP":
            raise H5Error(f"local heap signature missing at {heap_addr}")
        seg_size, seg = self._len(p + 8), self._addr(self._off(p + 8 + 2 * self.sl))
        if offset >= seg_size:
            raise H5Error("link name offset outside the local heap")
        self._need(seg, seg_size, "local heap data segment")
        raw = bytes(self.buf[seg + offset:seg + seg_size])
        end = raw.find(b"\0")
        if end < 0:
            raise H5Error("unterminated link name")
        try:
            return raw[:end].decode("utf-8")
        except UnicodeDecodeError:
            raise H5Error("link name is not UTF-8") from None

    def _group_entries(self, btree_addr, heap_addr, depth=0):
        """[(name, symbol entry)] in B-tree order (names ascending)."""
        if depth > 16:
            raise H5Error("group B-tree too deep (cycle?)")
        p = self._addr(btree_addr)
        self._need(p, 8 + 2 * self.so, "B-tree node")
        if bytes(self.buf[p:p + 4]) != b"TREE":
            raise H5Error(f"B-tree signature missing at {btree_addr}")
        ntype, level, used = self._u(p + 4, 1), self._u(p + 5, 1), self._u(p + 6, 2)
        if ntype != 0:
            raise H5Error("group B-tree expected (node type 0)")
        q = p + 8 + 2 * self.so                        # after left / right sibling
        out = []
        for i in range(used):
            child = self._off(q + self.sl + i * (self.sl + self.so))      # key_i, child_i, key_i+1, ...
            if level > 0:
                out += self._group_entries(child, heap_addr, depth + 1)
            else:
                s = self._addr(child)
                if bytes(self.buf[s:s + 4]) != b"SNOD":
                    raise H5Error(f"symbol table node signature missing at {child}")
                n = self._u(s + 6, 2)
                for k in range(n):
                    e = self._symbol_entry(s + 8 + k * self.entry_size)
                    out.append((self._heap_name(heap_addr, e["name_off"]), e))
        return out

    def _children(self, entry):
        """{name: symbol entry} if the object is a group, else None."""
        if entry.get("cache") == 1:
            return dict(self._group_entries(entry["btree"], entry["heap"]))
        for mtype, _, body in self._messages(entry["header"]):
            if mtype == 0x0011:                        # symbol table message: B-tree address, local heap address
                return dict(self._group_entries(int.from_bytes(body[:self.so], "little"),
                                                int.from_bytes(body[self.so:2 * self.so], "little")))
            if mtype in (0x0002, 0x0006):
                raise H5Error("link-info / link messages (new-style groups) are not supported")
        return None

    def walk(self):
        """yields (path, symbol entry, is_group) depth first, names ascending."""
        stack = [("", self.root)]
        seen = set()
        while stack:
            path, e = stack.pop()
            if e["header"] in seen:
                continue
            seen.add(e["header"])
            kids = self._children(e)
            yield path or "/", e, kids is not None
            if kids:
                for name in sorted(kids, reverse=True):
                    stack.append((f"{path}/{name}", kids[name]))

    def groups(self):
        return [p for p, _, g in self.walk() if g]

    def datasets(self):
        return {p.lstrip("/"): self._dataset(e["header"]) for p, e, g in self.walk() if not g}

    def __getitem__(self, path):
        e = self.root
        for part in [x for x in path.split("/") if x]:
            kids = self._children(e)
            if kids is None or part not in kids:
                raise KeyError(path)
            e = kids[part]
        if self._children(e) is not None:
            raise KeyError(f"{path} is a group")
        return self._dataset(e["header"])

    # ---- datasets ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _dtype(body):
        cls, version = body[0] & 0x0F, body[0] >> 4
        bits0 = body[1]
        size = int.from_bytes(body[4:8], "little")
        order = ">" if bits0 & 1 else "<"
        if version not in (1, 2, 3):
            raise H5Error(f"datatype message version {version}")
        if cls == 1:                                   # floating point: trust the size for the IEEE formats
            if size not in (2, 4, 8):
                raise H5Error(f"{size}-byte float")
            return np.dtype(f"{order}f{size}")
        if cls == 0:
            if size not in (1, 2, 4, 8):
                raise H5Error(f"{size}-byte integer")
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
        names = {2: "time", 3: "string", 4: "bitfield", 5: "opaque", 6: "compound", 7: "reference", 8: "enum", 9: "variable-length", 10: "array"}
        raise H5Error(f"datatype class {names.get(cls, cls)} is not supported")

    def _dataspace(self, body):
        version, rank, flags = body[0], body[1], body[2]
        if rank > 32:
            raise H5Error(f"dataspace rank {rank}")
        if version == 1:
            p = 8
        elif version == 2:
            p = 4
            if body[3] == 2:
                raise H5Error("null dataspace")
        else:
            raise H5Error(f"dataspace message version {version}")
        return tuple(int.from_bytes(body[p + i * self.sl:p + (i + 1) * self.sl], "little") for i in range(rank))

    def _filters(self, body):
        version, n = body[0], body[1]
        p = 8 if version == 1 else 2
        out = []
        for _ in range(n):
            fid = int.from_bytes(body[p:p + 2], "little")
            if version == 1 or fid >= 256:
                name_len = int.from_bytes(body[p + 2:p + 4], "little")
                p_vals = p + 8 + (((name_len + 7) // 8) * 8 if version == 1 else name_len)
            else:
                name_len, p_vals = 0, p + 6
            flags_pos, n_pos = (p + 4, p + 6) if (version == 1 or fid >= 256) else (p + 2, p + 4)
            nvals = int.from_bytes(body[n_pos:n_pos + 2], "little")
            vals = [int.from_bytes(body[p_vals + 4 * i:p_vals + 4 * i + 4], "little") for i in range(nvals)]
            p = p_vals + 4 * nvals
            if version == 1 and nvals % 2:
                p += 4
            if fid not in (1, 2):
                raise H5Error(f"filter {fid} is not supported (only deflate and shuffle)")
            out.append((fid, vals))
            del flags_pos
        return out

    def _dataset(self, header):
        shape = dtype = layout = None
        filters = []
        for mtype, _, body in self._messages(header):
            body = bytes(body)
            if mtype == 0x0001:
                shape = self._dataspace(body)
            elif mtype == 0x0003:
                dtype = self._dtype(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000B:
                filters = self._filters(body)
            elif mtype == 0x0007:
                raise H5Error("external data files are not supported")
        if shape is None or dtype is None or layout is None:
            raise H5Error("object is neither a group nor a complete dataset")
        count = 1
        for d in shape:
            count *= d
        nbytes = count * dtype.itemsize
        if nbytes > max(1 << 20, 1100 * len(self.buf)):      # deflate expands at most ~1030x
            raise H5Error(f"dataspace of {nbytes} bytes cannot be stored in a file of {len(self.buf)} bytes")
        version = layout[0]
        if version == 3:
            cls = layout[1]
            if cls == 0:
                size = int.from_bytes(layout[2:4], "little")
                raw = layout[4:4 + size]
            elif cls == 1:
                addr = int.from_bytes(layout[2:2 + self.so], "little")
                raw = self._contiguous(addr, nbytes)
            elif cls == 2:
                rank1 = layout[2]
                addr = int.from_bytes(layout[3:3 + self.so], "little")
                dims = [int.from_bytes(layout[3 + self.so + 4 * i:7 + self.so + 4 * i], "little") for i in range(rank1)]
                raw = self._chunked(addr, shape, dims[:-1], dtype, filters)
            else:
                raise H5Error(f"layout class {cls}")
        elif version in (1, 2):
            rank1, cls = layout[1], layout[2]
            p = 8
            addr = None
            if cls != 0:
                addr = int.from_bytes(layout[p:p + self.so], "little")
                p += self.so
            dims = [int.from_bytes(layout[p + 4 * i:p + 4 * i + 4], "little") for i in range(rank1)]
            p += 4 * rank1
            if cls == 1:
                raw = self._contiguous(addr, nbytes)
            elif cls == 2:
                raw = self._chunked(addr, shape, dims[:-1], dtype, filters)      # last "dimension" = element size
            else:
                size = int.from_bytes(layout[p:p + 4], "little")
                raw = layout[p + 4:p + 4 + size]
        else:
            raise H5Error(f"data layout message version {version}")
        if len(raw) < nbytes:
            raise H5Error("dataset storage shorter than its dataspace")
        return np.frombuffer(bytes(raw[:nbytes]), dtype=dtype).reshape(shape).astype(dtype.newbyteorder("="))

    def _contiguous(self, addr, nbytes):
        if addr == UNDEF >> (64 - 8 * self.so):
            return bytes(nbytes)                        # never written: fill value (0)
        p = self._addr(addr)
        self._need(p, nbytes, "contiguous dataset")
        return self.buf[p:p + nbytes]

    def _chunk_leaves(self, addr, rank, depth=0):
        if depth > 16:
            raise H5Error("chunk B-tree too deep")
        p = self._addr(addr)
        if bytes(self.buf[p:p + 4]) != b"TREE" or self._u(p + 4, 1) != 1:
            raise H5Error("raw-data chunk B-tree expected (node type 1)")
        level, used = self._u(p + 5, 1), self._u(p + 6, 2)
        key = 8 + 8 * (rank + 1)
        q = p + 8 + 2 * self.so
        for i in range(used):
            k = q + i * (key + self.so)
            size, mask = self._u(k, 4), self._u(k + 4, 4)
            offs = [self._u(k + 8 + 8 * d, 8) for d in range(rank)]
            child = self._off(k + key)
            if level > 0:
                yield from self._chunk_leaves(child, rank, depth + 1)
            else:
                yield size, mask, offs, child

    def _chunked(self, addr, shape, chunk, dtype, filters):
        rank = len(shape)
        if len(chunk) != rank or any(c <= 0 for c in chunk):
            raise H5Error("chunk shape does not fit the dataspace")
        chunk_bytes = dtype.itemsize
        for c in chunk:
            chunk_bytes *= c
        if chunk_bytes > max(1 << 20, 1100 * len(self.buf)):
            raise H5Error("chunk larger than the file can hold")
        out = np.zeros(shape, dtype=dtype)
        if addr == UNDEF >> (64 - 8 * self.so):
            return out.tobytes()
        for size, mask, offs, child in self._chunk_leaves(addr, rank):
            p = self._addr(child)
            self._need(p, size, "chunk")
            raw = bytes(self.buf[p:p + size])
            for i, (fid, _) in reversed(list(enumerate(filters))):
                if mask >> i & 1:
                    continue
                if fid == 1:
                    try:
                        raw = zlib.decompress(raw)
                    except zlib.error as e:
                        raise H5Error(f"deflate: {e}") from None
                else:                                   # shuffle: bytes of every element are stored plane by plane
                    n = len(raw) // dtype.itemsize
                    raw = np.frombuffer(raw[:n * dtype.itemsize], np.uint8).reshape(dtype.itemsize, n).T.tobytes()
            if len(raw) < chunk_bytes:
                raise H5Error("chunk shorter than its declared shape")
            if any(o >= s for o, s in zip(offs, shape)):
                raise H5Error("chunk offset outside the dataspace")
            block = np.frombuffer(raw[:chunk_bytes], dtype=dtype).reshape(chunk)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
            out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out.tobytes()


def read(path_or_bytes):
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        return File(path_or_bytes)
    with open(path_or_bytes, "rb") as f:
        return File(f.read())


# =====================================================================================================================
# writer (same subset; used by the tests and to export weights for a Keras-side comparison)
# =====================================================================================================================
class _Writer:
    def __init__(self, leaf_k=4, internal_k=16):
        self.buf = bytearray()
        self.leaf_k, self.internal_k = leaf_k, internal_k

    def alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def message(mtype, body, flags=0):
        body = bytes(body) + bytes(-len(body) % 8)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def object_header(self, messages, split=False):
        """version-1 header; split=True puts the last message into a continuation block (as h5py does when attributes
        are added later)."""
        if split and len(messages) > 1:
            tail = messages[-1]
            taddr = self.alloc(tail)
            messages = messages[:-1] + [self.message(0x0010, struct.pack("<QQ", taddr, len(tail)))]
            count = len(messages) + 1
        else:
            count = len(messages)
        body = b"".join(messages)
        return self.alloc(struct.pack("<BxHII4x", 1, count, 1, len(body)) + body)

    def dataset(self, arr, chunks=None, deflate=False, shuffle=False, layout_version=3, split=False):
        arr = np.asarray(arr)
        if not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)
        dt = arr.dtype
        if dt.kind == "f":
            m = {2: (10, 5, 15, 0, 10), 4: (23, 8, 31, 0, 23), 8: (52, 11, 63, 0, 52)}[dt.itemsize]
            bias = {2: 15, 4: 127, 8: 1023}[dt.itemsize]
            dtype_msg = struct.pack("<BBBBI", 0x11, 0x20, m[2], 0, dt.itemsize) + \
                struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, m[0], m[1], 0, m[0], bias)
        elif dt.kind in "iu":
            dtype_msg = struct.pack("<BBBBI", 0x10, 8 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
        else:
            raise H5Error(f"cannot write dtype {dt}")
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        msgs = [self.message(0x0001, space), self.message(0x0003, dtype_msg, flags=1)]
        if chunks is None:
            addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
            if layout_version == 3:
                layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
            else:
                layout = struct.pack("<BBB5xQ", layout_version, arr.ndim, 1, addr) + b"".join(struct.pack("<I", d) for d in arr.shape)
        else:
            filters = ([(2, [dt.itemsize])] if shuffle else []) + ([(1, [6])] if deflate else [])
            if filters:
                fb = struct.pack("<BB6x", 1, len(filters))
                for fid, vals in filters:
                    fb += struct.pack("<HHHH", fid, 0, 0, len(vals)) + b"".join(struct.pack("<I", v) for v in vals)
                    if len(vals) % 2:
                        fb += bytes(4)
                msgs.append(self.message(0x000B, fb))
            leaves = []
            grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
            for offs in np.ndindex(*[len(g) for g in grid]):
                o = [g[i] for g, i in zip(grid, offs)]
                block = np.zeros(chunks, dt)
                sel = tuple(slice(a, min(a + c, s)) for a, c, s in zip(o, chunks, arr.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sel)] = arr[sel]
                raw = block.tobytes()
                if shuffle:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, dt.itemsize).T.tobytes()
                if deflate:
                    raw = zlib.compress(raw, 6)
                leaves.append((len(raw), o, self.alloc(raw)))
            bt = self._chunk_btree(leaves, arr.ndim, list(arr.shape))
            layout = struct.pack("<BBBQ", 3, 2, arr.ndim + 1, bt) + b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", dt.itemsize)
        msgs.append(self.message(0x0008, layout))
        return self.object_header(msgs, split=split)

    def _chunk_btree(self, leaves, rank, shape):
        def node(level, entries, last_key):
            body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(entries), UNDEF, UNDEF)
            for size, offs, child in entries:
                body += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0) + struct.pack("<Q", child)
            body += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", o) for o in last_key) + struct.pack("<Q", 0)
            return self.alloc(body)
        cap = 2 * self.internal_k
        level = 0
        nodes = leaves
        while True:
            parents = []
            for i in range(0, max(len(nodes), 1), cap):
                part = nodes[i:i + cap]
                addr = node(level, part, shape)
                parents.append((0, part[0][1] if part else [0] * rank, addr))
            if len(parents) == 1:
                return parents[0][2]
            nodes, level = parents, level + 1

    def group(self, children):
        """children: {name: object header address} -> (header address, btree, heap)"""
        names = sorted(children)
        heap_data = bytearray(b"\0" * 8)               # offset 0 = the empty name (first key of the B-tree)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += n.encode("utf-8") + b"\0"
            heap_data += bytes(-len(heap_data) % 8)
        seg = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, seg))
        cap = 2 * self.leaf_k
        snods = []
        for i in range(0, len(names), cap):
            part = names[i:i + cap]
            body = b"SNOD" + struct.pack("<BxH", 1, len(part))
            for n in part:
                body += struct.pack("<QQII16x", offs[n], children[n][0], 1 if children[n][1] else 0, 0)[:24]
                if children[n][1]:
                    body += struct.pack("<QQ", *children[n][1])
                else:
                    body += bytes(16)
            body += bytes((cap - len(part)) * 40)
            snods.append((offs[part[-1]], self.alloc(body)))
        level, nodes, icap = 0, snods, 2 * self.internal_k
        while True:
            parents = []
            for i in range(0, max(len(nodes), 1), icap):
                part = nodes[i:i + icap]
                body = b"TREE" + struct.pack("<BBHQQ", 0, level, len(part), UNDEF, UNDEF) + struct.pack("<Q", 0)
                for key, child in part:
                    body += struct.pack("<QQ", child, key)
                parents.append((part[-1][0] if part else 0, self.alloc(body)))
            if len(parents) == 1:
                btree = parents[0][1]
                break
            nodes, level = parents, level + 1
        header = self.object_header([self.message(0x0011, struct.pack("<QQ", btree, heap))])
        return header, btree, heap


def write(tree, chunks=None, deflate=False, shuffle=False, leaf_k=4, internal_k=16, layout_version=3, split_headers=False):
    """tree: nested dict {name: array | dict} -> bytes of an HDF5 file (superblock v0, old-style groups).
    chunks: None (contiguous) or a function array -> chunk shape | None."""
    w = _Writer(leaf_k, internal_k)
    w.buf += bytes(96)                                 # superblock v0 with 8-byte offsets: 56 + 40 bytes

    def emit(node):
        if isinstance(node, dict):
            kids = {}
            for name, child in node.items():
                kids[name] = emit(child)
            header, btree, heap = w.group(kids)
            return header, (btree, heap)
        ch = chunks(node) if callable(chunks) else chunks
        return w.dataset(np.asarray(node), chunks=ch, deflate=deflate and ch is not None, shuffle=shuffle and ch is not None,
                         layout_version=layout_version, split=split_headers), None

    header, (btree, heap) = emit(tree)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, internal_k, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(w.buf), UNDEF)
    sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    w.buf[:96] = sb
    return bytes(w.buf)
