"""Minimal HDF5 reader/writer, just enough for Keras weight files.

The reference saves and loads its network with ``keras.Model.save`` / ``keras.models.load_model`` (``network.py:217-253``;
``train.py:43-45,172``; ``usi.py:26-29``), i.e. HDF5 files written by h5py, which this environment does not have.
This module restates the parts of the published HDF5 file format (HDF5 File Format Specification, version 0/1
structures, which is what h5py writes by default) those files use:

* reader: superblock v0/v1 (and v2/v3 with compact link messages), object headers v1 and v2, old-style groups
  (symbol-table message, B-tree v1, local heap, symbol nodes), datasets with contiguous, compact or unfiltered
  chunked layout, fixed-point / floating-point / fixed-length string / variable-length string datatypes, attributes
  (message versions 1-3).  Compressed datasets and dense (fractal-heap) groups raise ``NotImplementedError``.
* writer: old-style groups, contiguous little-endian datasets, fixed-length string attributes -- the layout of
  ``keras.Model.save_weights``.

The reader is pinned against a file written by the real HDF5 library (a MATLAB v7.3 file in scipy's test data; see
``tests/test_hdf5.py``); the writer is checked through the reader.
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(n):
    return (n + 7) & ~7


# =====================================================================================================================
# reader
# =====================================================================================================================
class _Datatype:
    def __init__(self, cls, size, dtype=None, vlen_string=False, strpad=0):
        self.cls, self.size, self.dtype, self.vlen_string, self.strpad = cls, size, dtype, vlen_string, strpad


class _Object:
    """A parsed object header: its messages by type."""

    def __init__(self, f, addr):
        self.file, self.addr = f, addr
        self.messages = f._read_header(addr)

    def first(self, mtype):
        for t, data in self.messages:
            if t == mtype:
                return data
        return None

    @property
    def attrs(self):
        out = {}
        for t, data in self.messages:
            if t == 0x000C:
                name, value = self.file._parse_attribute(data)
                out[name] = value
        if self.first(0x0015) is not None and not out:
            info = self.first(0x0015)
            if len(info) >= 2 and struct.unpack_from("<Q", info, 2 + (2 if info[1] & 1 else 0))[0] != UNDEF:
                raise NotImplementedError("densely stored attributes (fractal heap) are not supported")
        return out


class Dataset(_Object):
    def __init__(self, f, addr):
        super().__init__(f, addr)
        self.shape = f._parse_dataspace(self.first(0x0001))
        self._dt = f._parse_datatype(self.first(0x0003))
        self.dtype = self._dt.dtype

    def read(self):
        f = self.file
        if self.first(0x000B) is not None and f._has_filters(self.first(0x000B)):
            raise NotImplementedError("filtered (compressed) datasets are not supported")
        if self._dt.dtype is None:
            raise NotImplementedError(f"datatype class {self._dt.cls} is not supported")
        count = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        raw = f._read_layout(self.first(0x0008), self.shape, self._dt.size, count)
        return f._decode(raw, self._dt, self.shape)

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a


class Group(_Object):
    def __init__(self, f, addr):
        super().__init__(f, addr)
        self._links = None

    def _load(self):
        if self._links is not None:
            return self._links
        f, links = self.file, {}
        st = self.first(0x0011)
        if st is not None:                                   # old-style group: B-tree v1 + local heap
            btree, heap = struct.unpack_from("<QQ", st)
            heap_data = f._local_heap(heap)
            for name_off, obj_addr in f._walk_group_btree(btree):
                end = heap_data.index(b"\0", name_off)
                links[heap_data[name_off:end].decode("utf-8")] = obj_addr
        else:                                                # new-style group with compact link messages
            info = self.first(0x0002)
            if info is not None:
                flags = info[1]
                off = 2 + (8 if flags & 1 else 0)
                if struct.unpack_from("<Q", info, off)[0] != UNDEF:
                    raise NotImplementedError("densely stored groups (fractal heap) are not supported")
            for t, data in self.messages:
                if t == 0x0006:
                    name, addr = f._parse_link(data)
                    if addr is not None:
                        links[name] = addr
        self._links = links
        return links

    def keys(self):
        return list(self._load().keys())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError(path)
            node = node.file._open(links[part])
        return node


class File(Group):
    """``File(path)`` or ``File(bytes)``; indexing and ``attrs`` follow the h5py conventions the Keras code uses."""

    def __init__(self, source):
        if isinstance(source, (bytes, bytearray, memoryview)):
            self.buf = bytes(source)
        else:
            with open(source, "rb") as fh:
                self.buf = fh.read()
        self._cache = {}
        sb = 0
        while self.buf[sb:sb + 8] != SIGNATURE:              # the superblock sits at 0, 512, 1024, ... (user block)
            sb = 512 if sb == 0 else sb * 2
            if sb + 8 > len(self.buf):
                raise ValueError("not an HDF5 file")
        version = self.buf[sb + 8]
        if version in (0, 1):
            size_off, size_len = self.buf[sb + 13], self.buf[sb + 14]
            if (size_off, size_len) != (8, 8):
                raise NotImplementedError("only 8-byte offsets and lengths are supported")
            p = sb + 24 + (4 if version == 1 else 0)
            self.base, _free, _eof, _drv = struct.unpack_from("<QQQQ", self.buf, p)
            root_addr = struct.unpack_from("<Q", self.buf, p + 32 + 8)[0]      # symbol table entry: name offset, header address
        elif version in (2, 3):
            if (self.buf[sb + 9], self.buf[sb + 10]) != (8, 8):
                raise NotImplementedError("only 8-byte offsets and lengths are supported")
            self.base, _ext, _eof, root_addr = struct.unpack_from("<QQQQ", self.buf, sb + 12)
        else:
            raise NotImplementedError(f"superblock version {version}")
        if self.base == UNDEF:
            self.base = 0
        if self.base == 0 and sb != 0:
            self.base = sb                                   # addresses are relative to the superblock's position
        Group.__init__(self, self, root_addr)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    # ---- low-level pieces -------------------------------------------------------------------------------------
    def _at(self, addr):
        return self.base + addr

    def _open(self, addr):
        if addr not in self._cache:
            probe = _Object(self, addr)
            is_dataset = probe.first(0x0008) is not None and probe.first(0x0003) is not None
            self._cache[addr] = Dataset(self, addr) if is_dataset else Group(self, addr)
        return self._cache[addr]

    def _read_header(self, addr):
        buf, p = self.buf, self._at(addr)
        messages = []
        if buf[p:p + 4] == b"OHDR":                           # version 2 object header
            flags = buf[p + 5]
            q = p + 6
            if flags & 0x20:
                q += 16
            if flags & 0x10:
                q += 4
            width = 1 << (flags & 3)
            chunk_size = int.from_bytes(buf[q:q + width], "little")
            q += width
            blocks = [(q, q + chunk_size)]
            while blocks:
                q, end = blocks.pop(0)
                while q + 4 <= end:
                    mtype, size, _mflags = buf[q], struct.unpack_from("<H", buf, q + 1)[0], buf[q + 3]
                    q += 4 + (2 if flags & 0x04 else 0)
                    data = buf[q:q + size]
                    q += size
                    if mtype == 0x10:
                        off, length = struct.unpack_from("<QQ", data)
                        blocks.append((self._at(off) + 4, self._at(off) + length - 4))   # skip "OCHK", stop before the checksum
                    elif mtype != 0:
                        messages.append((mtype, data))
            return messages
        version, _res, nmsgs, _refs, header_size = struct.unpack_from("<BBHII", buf, p)
        if version != 1:
            raise NotImplementedError(f"object header version {version} at {addr:#x}")
        blocks = [(p + 16, p + 16 + header_size)]
        while blocks and len(messages) < nmsgs:
            q, end = blocks.pop(0)
            while q + 8 <= end and len(messages) < nmsgs:
                mtype, size, _mflags = struct.unpack_from("<HHB", buf, q)
                data = buf[q + 8:q + 8 + size]
                q += 8 + size
                if mtype == 0x0010:
                    off, length = struct.unpack_from("<QQ", data)
                    blocks.append((self._at(off), self._at(off) + length))
                messages.append((mtype, data))
        return [(t, d) for t, d in messages if t not in (0x0000, 0x0010)]

    def _local_heap(self, addr):
        p = self._at(addr)
        if self.buf[p:p + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        size, _free, data_addr = struct.unpack_from("<QQQ", self.buf, p + 8)
        q = self._at(data_addr)
        return self.buf[q:q + size]

    def _walk_group_btree(self, addr):
        p = self._at(addr)
        if self.buf[p:p + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        node_type, level, used = struct.unpack_from("<BBH", self.buf, p + 4)
        if node_type != 0:
            raise ValueError("not a group B-tree")
        q = p + 24
        for i in range(used):
            child = struct.unpack_from("<Q", self.buf, q + 8 + i * 16)[0]      # key, child, key, child, ..., key
            if level > 0:
                yield from self._walk_group_btree(child)
            else:
                s = self._at(child)
                if self.buf[s:s + 4] != b"SNOD":
                    raise ValueError("bad symbol node signature")
                n = struct.unpack_from("<H", self.buf, s + 6)[0]
                for k in range(n):
                    yield struct.unpack_from("<QQ", self.buf, s + 8 + k * 40)

    def _parse_link(self, data):
        flags = data[1]
        q = 2
        link_type = 0
        if flags & 0x08:
            link_type = data[q]
            q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        width = 1 << (flags & 3)
        n = int.from_bytes(data[q:q + width], "little")
        q += width
        name = data[q:q + n].decode("utf-8")
        q += n
        return name, (struct.unpack_from("<Q", data, q)[0] if link_type == 0 else None)

    @staticmethod
    def _parse_dataspace(data):
        version, rank, flags = data[0], data[1], data[2]
        if version == 1:
            q = 8
        elif version == 2:
            if data[3] == 2:                                  # null dataspace
                return (0,)
            q = 4
        else:
            raise NotImplementedError(f"dataspace version {version}")
        return tuple(struct.unpack_from(f"<{rank}Q", data, q)) if rank else ()

    def _parse_datatype(self, data):
        cls, bits0 = data[0] & 0x0F, data[1]
        size = struct.unpack_from("<I", data, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return _Datatype(cls, size, np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}"))
        if cls == 1:
            return _Datatype(cls, size, np.dtype(f"{order}f{size}"))
        if cls == 3:
            return _Datatype(cls, size, np.dtype(f"S{size}"), strpad=bits0 & 0x0F)
        if cls == 9:
            is_string = (bits0 & 0x0F) == 1
            return _Datatype(cls, size, np.dtype(object) if is_string else None, vlen_string=is_string)
        return _Datatype(cls, size, None)

    @staticmethod
    def _has_filters(data):
        return data[1] > 0

    def _read_layout(self, data, shape, elem, count):
        version = data[0]
        nbytes = count * elem
        if version == 3:
            cls = data[1]
            if cls == 0:
                size = struct.unpack_from("<H", data, 2)[0]
                return data[4:4 + size]
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", data, 2)
                if addr == UNDEF:
                    return b"\0" * nbytes
                return self.buf[self._at(addr):self._at(addr) + nbytes]
            if cls == 2:
                ndims = data[2]
                btree = struct.unpack_from("<Q", data, 3)[0]
                chunk = struct.unpack_from(f"<{ndims}I", data, 11)
                return self._read_chunked(btree, shape, chunk[:-1], elem)
            raise NotImplementedError(f"data layout class {cls}")
        if version in (1, 2):
            ndims, cls = data[1], data[2]
            q = 8
            if cls == 1:
                addr = struct.unpack_from("<Q", data, q)[0]
                return self.buf[self._at(addr):self._at(addr) + nbytes]
            if cls == 0:
                q += 4 * ndims
                size = struct.unpack_from("<I", data, q)[0]
                return data[q + 4:q + 4 + size]
            raise NotImplementedError("chunked layout version 1/2")
        raise NotImplementedError(f"data layout version {version}")

    def _read_chunked(self, btree, shape, chunk, elem):
        out = np.zeros(shape, dtype=f"V{elem}")
        if btree == UNDEF:
            return out.tobytes()
        rank = len(shape)

        def walk(addr):
            p = self._at(addr)
            if self.buf[p:p + 4] != b"TREE":
                raise ValueError("bad chunk B-tree signature")
            _ntype, level, used = struct.unpack_from("<BBH", self.buf, p + 4)
            key_size = 8 + 8 * (rank + 1)
            q = p + 24
            for i in range(used):
                k = q + i * (key_size + 8)
                nbytes, mask = struct.unpack_from("<II", self.buf, k)
                offs = struct.unpack_from(f"<{rank}Q", self.buf, k + 8)
                child = struct.unpack_from("<Q", self.buf, k + key_size)[0]
                if level > 0:
                    walk(child)
                    continue
                if mask:
                    raise NotImplementedError("filtered chunks are not supported")
                block = np.frombuffer(self.buf, dtype=f"V{elem}", count=int(np.prod(chunk)), offset=self._at(child)).reshape(chunk)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(btree)
        return out.tobytes()

    def _decode(self, raw, dt, shape):
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if dt.vlen_string:
            vals = []
            for i in range(count):
                length, coll, index = struct.unpack_from("<IQI", raw, i * 16)
                vals.append(self._global_heap_object(coll, index)[:length].decode("utf-8", "replace") if length else "")
            arr = np.array(vals, dtype=object)
        else:
            arr = np.frombuffer(raw, dtype=dt.dtype, count=count).copy()
            if dt.cls in (0, 1) and arr.dtype.byteorder == ">":
                arr = arr.astype(arr.dtype.newbyteorder("<"))
        return arr.reshape(shape) if shape else arr.reshape(())

    def _global_heap_object(self, coll_addr, index):
        p = self._at(coll_addr)
        if self.buf[p:p + 4] != b"GCOL":
            raise ValueError("bad global heap signature")
        size = struct.unpack_from("<Q", self.buf, p + 8)[0]
        q, end = p + 16, p + size
        while q + 16 <= end:
            idx, _refs, _res, osize = struct.unpack_from("<HHIQ", self.buf, q)
            if idx == 0:
                break
            if idx == index:
                return self.buf[q + 16:q + 16 + osize]
            q += 16 + _pad8(osize)
        raise KeyError(f"global heap object {index}")

    def _parse_attribute(self, data):
        version = data[0]
        if version == 1:
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", data, 2)
            q = 8
            name = data[q:q + name_size].split(b"\0")[0].decode("utf-8")
            q += _pad8(name_size)
            dt = self._parse_datatype(data[q:q + dt_size])
            q += _pad8(dt_size)
            shape = self._parse_dataspace(data[q:q + ds_size])
            q += _pad8(ds_size)
        elif version in (2, 3):
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", data, 2)
            q = 8 + (1 if version == 3 else 0)
            name = data[q:q + name_size].split(b"\0")[0].decode("utf-8")
            q += name_size
            dt = self._parse_datatype(data[q:q + dt_size])
            q += dt_size
            shape = self._parse_dataspace(data[q:q + ds_size])
            q += ds_size
        else:
            raise NotImplementedError(f"attribute message version {version}")
        if dt.dtype is None:
            return name, None
        value = self._decode(data[q:], dt, shape)
        if value.shape == () and not dt.vlen_string:
            value = value[()]
        elif value.shape == ():
            value = value[()]
        return name, value


# =====================================================================================================================
# writer (old-style groups, contiguous datasets, fixed-length string attributes)
# =====================================================================================================================
class _WGroup:
    def __init__(self):
        self.children = {}     # name -> _WGroup | np.ndarray
        self.attrs = {}        # name -> np.ndarray of dtype S* / float / int, or bytes

    def require_group(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            node = node.children.setdefault(part, _WGroup())
            if not isinstance(node, _WGroup):
                raise ValueError(f"{part} is a dataset")
        return node

    def create_dataset(self, path, data):
        parts = [p for p in path.split("/") if p]
        parent = self.require_group("/".join(parts[:-1]))
        parent.children[parts[-1]] = np.ascontiguousarray(data)


class Writer(_WGroup):
    """Collects groups / datasets / attributes, then :meth:`tobytes` lays the file out."""

    GROUP_LEAF_K, GROUP_INTERNAL_K = 4, 16

    def tobytes(self):
        self._buf = bytearray(b"\0" * 96)                    # superblock (56 bytes + root symbol table entry 40 bytes)
        root_header, root_btree, root_heap = self._write_group(self)
        sb = bytearray()
        sb += SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0])
        sb += struct.pack("<HHI", self.GROUP_LEAF_K, self.GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self._buf), UNDEF)
        sb += struct.pack("<QQII", 0, root_header, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
        self._buf[0:96] = sb
        return bytes(self._buf)

    def save(self, path):
        with open(path, "wb") as fh:
            fh.write(self.tobytes())

    # ---- pieces ---------------------------------------------------------------------------------------------------
    def _alloc(self, data):
        while len(self._buf) % 8:
            self._buf += b"\0"
        addr = len(self._buf)
        self._buf += data
        return addr

    @staticmethod
    def _datatype(arr):
        if arr.dtype.kind == "f":
            size = arr.dtype.itemsize
            exp_bits, mant_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
            return (struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size) +
                    struct.pack("<HHBBBBI", 0, size * 8, mant_bits, exp_bits, 0, mant_bits, bias))
        if arr.dtype.kind in "iu":
            size = arr.dtype.itemsize
            return struct.pack("<BBBBI", 0x10, 0x08 if arr.dtype.kind == "i" else 0x00, 0, 0, size) + struct.pack("<HH", 0, size * 8)
        if arr.dtype.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, arr.dtype.itemsize)      # null-padded ASCII
        raise TypeError(f"cannot store dtype {arr.dtype}")

    @staticmethod
    def _dataspace(shape):
        return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)

    @staticmethod
    def _message(mtype, data, flags=0):
        data = data + b"\0" * (_pad8(len(data)) - len(data))
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    def _attribute(self, name, value):
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.dtype.kind == "S" and arr.dtype.itemsize == 0:
            arr = arr.astype("S1")
        nm = name.encode("utf-8") + b"\0"
        dt, ds = self._datatype(arr), self._dataspace(arr.shape)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
        body += nm + b"\0" * (_pad8(len(nm)) - len(nm))
        body += dt + b"\0" * (_pad8(len(dt)) - len(dt))
        body += ds + b"\0" * (_pad8(len(ds)) - len(ds))
        body += arr.tobytes()
        if len(body) > 0xFFF0:
            raise ValueError(f"attribute {name} is too large for an object header message")
        return self._message(0x000C, body)

    def _object_header(self, messages):
        body = b"".join(messages)
        return self._alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)

    def _write_dataset(self, arr):
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        data_addr = self._alloc(arr.tobytes()) if arr.size else UNDEF
        msgs = [self._message(0x0001, self._dataspace(arr.shape)),
                self._message(0x0003, self._datatype(arr), flags=1),
                self._message(0x0005, struct.pack("<BBB", 2, 0x02, 0x02)),       # fill value v2: late allocation, never written
                self._message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes))]
        return self._object_header(msgs)

    def _write_group(self, group):
        entries = []                                          # (name, object header address, cache: None | (btree, heap))
        for name in sorted(group.children):
            child = group.children[name]
            if isinstance(child, _WGroup):
                header, btree, heap = self._write_group(child)
                entries.append((name, header, (btree, heap)))
            else:
                entries.append((name, self._write_dataset(child), None))
        # local heap: "" at offset 0 (the B-tree's first key), then the names
        heap_data, offsets = bytearray(b"\0" * 8), {}
        for name, _addr, _cache in entries:
            offsets[name] = len(heap_data)
            raw = name.encode("utf-8") + b"\0"
            heap_data += raw + b"\0" * (_pad8(len(raw)) - len(raw))
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16) if True else b""   # one free block (next = 1: last, size 16) closing the segment
        data_addr = self._alloc(bytes(heap_data))
        heap_addr = self._alloc(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap_data), free_off, data_addr))
        # symbol nodes of up to 2K entries, sorted by name; one B-tree leaf level above them
        per = 2 * self.GROUP_LEAF_K
        nodes, keys = [], [0]
        for i in range(0, len(entries), per):
            chunk = entries[i:i + per]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for name, addr, cache in chunk:
                body += struct.pack("<QQII", offsets[name], addr, 1 if cache else 0, 0)
                body += struct.pack("<QQ", *cache) if cache else b"\0" * 16
            body += b"\0" * (40 * (per - len(chunk)))
            nodes.append(self._alloc(body))
            keys.append(offsets[chunk[-1][0]])
        if len(nodes) > 2 * self.GROUP_INTERNAL_K:
            raise ValueError("too many entries for a single-level group B-tree")
        tree = b"TREE" + struct.pack("<BBH", 0, 0, len(nodes)) + struct.pack("<QQ", UNDEF, UNDEF)
        for i, node in enumerate(nodes):
            tree += struct.pack("<QQ", keys[i], node)
        tree += struct.pack("<Q", keys[len(nodes)])
        tree += b"\0" * (16 * (2 * self.GROUP_INTERNAL_K - len(nodes)))
        btree_addr = self._alloc(tree)
        msgs = [self._message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [self._attribute(k, v) for k, v in group.attrs.items()]
        return self._object_header(msgs), btree_addr, heap_addr


# =====================================================================================================================
# Keras weight files
# =====================================================================================================================
def read_keras_weights(source):
    """The ``get_weights()`` list of a file written by ``keras.Model.save`` (weights under ``/model_weights``, as the
    reference's ``Network.save`` produces, network.py:245-253) or ``save_weights`` (weights at the root): for every
    layer in ``layer_names`` order, every tensor in its ``weight_names`` order -- the order Keras itself restores them in."""
    f = File(source)
    g = f["model_weights"] if "model_weights" in f else f
    attrs = g.attrs

    def names(a, key):
        if key in a:
            return [n.decode("utf-8") if isinstance(n, bytes) else str(n) for n in np.atleast_1d(a[key])]
        out, i = [], 0                                        # Keras splits oversized attributes into key0, key1, ...
        while f"{key}{i}" in a:
            out += [n.decode("utf-8") if isinstance(n, bytes) else str(n) for n in np.atleast_1d(a[f"{key}{i}"])]
            i += 1
        if not out and key not in a and i == 0:
            raise KeyError(key)
        return out

    try:
        layer_names = names(attrs, "layer_names")
    except KeyError:
        raise ValueError("no layer_names attribute: not a Keras weight file") from None
    weights = []
    for layer in layer_names:
        lg = g[layer]
        try:
            weight_names = names(lg.attrs, "weight_names")
        except KeyError:
            weight_names = []
        for wn in weight_names:
            weights.append(np.asarray(lg[wn].read(), dtype=np.float32))
    return weights


def write_keras_weights(path, weights, layers):
    """Write ``weights`` (the ``get_weights()`` list) in the ``save_weights`` layout; ``layers`` is a list of
    ``(layer_name, [weight_name, ...])`` covering the list in order, e.g. ``("conv2d", ["conv2d/kernel:0", "conv2d/bias:0"])``."""
    w = Writer()
    w.attrs["layer_names"] = np.array([name.encode("utf-8") for name, _ in layers])
    w.attrs["backend"] = np.bytes_(b"tensorflow")
    w.attrs["keras_version"] = np.bytes_(b"2.2.4-tf")
    i = 0
    for name, weight_names in layers:
        g = w.require_group(name)
        g.attrs["weight_names"] = np.array([wn.encode("utf-8") for wn in weight_names]) if weight_names else np.zeros((0,), "S1")
        for wn in weight_names:
            g.create_dataset(wn, np.asarray(weights[i], dtype=np.float32))
            i += 1
    if i != len(weights):
        raise ValueError(f"layers describe {i} tensors, got {len(weights)}")
    w.save(path)
