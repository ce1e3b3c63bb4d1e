"""A writer for uncompressed ``.npz`` archives that does not funnel a large table through
one thread.

``numpy.savez`` (what ``Blobs.save_archive`` of the reference calls, ``magmap/cv/detector.py:
269-323``) copies every array to ``bytes`` in 16 MB pieces and hands each piece to ``zipfile``,
which checksums it and appends it with ``write()``: three serial passes over the data, the last
one under the file's inode lock.  For the blob table of a whole stack (17 MB at config 2, 132 MB
for eight of them, 1 GB at config 3) that is the longest host-side stage of
``detect_blobs_stack``.  An archive's layout is known before a byte is written (stored
entries: header and data sizes are fixed), so this writer sizes the file once, maps it, and
lets a few threads copy the large arrays into the mapping side by side while another thread
computes their CRC-32; small entries, pickled objects and the directory are written by the
caller's thread.  The file is an ordinary ZIP64-capable archive with ``name.npy`` members in
``numpy.lib.format`` - ``numpy.load`` reads it like one written by ``numpy.savez``.
"""
from __future__ import annotations

import io
import mmap
import os
import struct
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional

import numpy as np

#: arrays of at least this many bytes are copied by the thread pool
LARGE_BYTES = 4 << 20
#: the thresholds ``zipfile`` uses for the ZIP64 forms of the directory records
ZIP64_LIMIT = (1 << 31) - 1
ZIP_FILECOUNT_LIMIT = (1 << 16) - 1
_THREADS = 4
_PIECE = 2 << 20

_pool: List[Optional[ThreadPoolExecutor]] = [None]


def _executor() -> ThreadPoolExecutor:
    if _pool[0] is None:
        _pool[0] = ThreadPoolExecutor(max_workers=_THREADS + 1, thread_name_prefix="npz")
    return _pool[0]


class _Entry:
    __slots__ = ("name", "flags", "head", "data", "size", "crc", "offset")

    def __init__(self, key: str, val):
        fname = key + ".npy"
        try:
            self.name, self.flags = fname.encode("ascii"), 0
        except UnicodeEncodeError:
            self.name, self.flags = fname.encode("utf-8"), 0x800
        arr = np.asanyarray(val)
        self.data = None
        if (arr.flags.c_contiguous and not arr.dtype.hasobject and arr.nbytes >= LARGE_BYTES
                and arr.dtype.isnative):
            try:
                buf = io.BytesIO()
                np.lib.format.write_array_header_1_0(
                    buf, np.lib.format.header_data_from_array_1_0(arr))
                self.data = memoryview(arr).cast("B")
                self.head = buf.getvalue()
            except (ValueError, TypeError):
                # a header too long for format 1.0, or a dtype without the buffer protocol
                self.data = None
        if self.data is None:
            # small, pickled or strided: numpy's own serialisation, in memory
            buf = io.BytesIO()
            np.lib.format.write_array(buf, arr, allow_pickle=True)
            self.head = buf.getvalue()
        self.size = len(self.head) + (len(self.data) if self.data is not None else 0)
        self.crc = 0
        self.offset = 0

    def local_header(self, dos_time: int, dos_date: int) -> bytes:
        # sizes always in the ZIP64 extra field, as numpy's force_zip64 members have them
        extra = struct.pack("<HHQQ", 1, 16, self.size, self.size)
        return struct.pack("<4s5H3L2H", b"PK\x03\x04", 45, self.flags, 0, dos_time, dos_date,
                           self.crc, 0xFFFFFFFF, 0xFFFFFFFF, len(self.name),
                           len(extra)) + self.name + extra

    def central_header(self, dos_time: int, dos_date: int) -> bytes:
        big = []
        size = offset = None
        if self.size > ZIP64_LIMIT:
            big += [self.size, self.size]
            size = 0xFFFFFFFF
        if self.offset > ZIP64_LIMIT:
            big.append(self.offset)
            offset = 0xFFFFFFFF
        extra = struct.pack("<HH" + "Q" * len(big), 1, 8 * len(big), *big) if big else b""
        version = 45 if big else 20
        return struct.pack(
            "<4s4B4H3L5H2L", b"PK\x01\x02", 45, 3, version, 0, self.flags, 0, dos_time, dos_date,
            self.crc, self.size if size is None else size, self.size if size is None else size,
            len(self.name), len(extra), 0, 0, 0, 0o600 << 16,
            self.offset if offset is None else offset) + self.name + extra


def _end_records(count: int, cd_size: int, cd_offset: int) -> bytes:
    out = b""
    if count > ZIP_FILECOUNT_LIMIT or cd_offset > ZIP64_LIMIT or cd_size > ZIP64_LIMIT:
        out += struct.pack("<4sQ2H2L4Q", b"PK\x06\x06", 44, 45, 45, 0, 0, count, count, cd_size,
                           cd_offset)
        out += struct.pack("<4sLQL", b"PK\x06\x07", 0, cd_offset + cd_size, 1)
        count = min(count, 0xFFFF)
        cd_size = min(cd_size, 0xFFFFFFFF)
        cd_offset = min(cd_offset, 0xFFFFFFFF)
    return out + struct.pack("<4s4H2LH", b"PK\x05\x06", 0, 0, count, count, cd_size, cd_offset, 0)


def _crc_of(parts) -> int:
    crc = 0
    for p in parts:
        crc = zlib.crc32(p, crc)
    return crc


def savez(path, arrays: Dict[str, object], add_suffix: bool = True) -> None:
    """Write ``arrays`` to the uncompressed archive ``path``.  With ``add_suffix`` a missing
    ``.npz`` is appended, as ``numpy.savez`` does for a file name; without, the file is written
    under exactly that name, as ``numpy.savez`` does for an open file.  Values are anything
    ``numpy.asanyarray`` takes; objects are pickled as ``numpy.savez`` pickles them."""
    path = os.fspath(path)
    if add_suffix and not path.endswith(".npz"):
        path += ".npz"
    entries = [_Entry(k, v) for k, v in arrays.items()]
    now = time.localtime()[:6]
    dos_date = (max(now[0], 1980) - 1980) << 9 | now[1] << 5 | now[2]
    dos_time = now[3] << 11 | now[4] << 5 | now[5] // 2

    # layout: [local header, .npy header, data] per member, directory, end records
    pos = 0
    for e in entries:
        e.offset = pos
        pos += len(e.local_header(dos_time, dos_date)) + e.size
    cd_offset = pos
    ex = _executor()

    # checksums of the large members run beside the copies (zlib releases the GIL)
    crc_jobs = {id(e): ex.submit(_crc_of, (e.head, e.data)) for e in entries if e.data is not None}
    for e in entries:
        if e.data is None:
            e.crc = zlib.crc32(e.head)

    def finish_crcs():
        for e in entries:
            if e.data is not None:
                e.crc = crc_jobs[id(e)].result()

    def directory() -> bytes:
        cd = b"".join(e.central_header(dos_time, dos_date) for e in entries)
        return cd + _end_records(len(entries), len(cd), cd_offset)

    large = any(e.data is not None for e in entries)
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o666)
    try:
        mapped = None
        if large:
            # the directory's size does not depend on the checksums
            total = cd_offset + len(directory())
            try:
                os.posix_fallocate(fd, 0, total)     # no SIGBUS on a full disk later on
                mapped = mmap.mmap(fd, total)
            except (OSError, ValueError):
                mapped = None
                os.ftruncate(fd, 0)
        if mapped is not None:
            dst = np.frombuffer(mapped, dtype=np.uint8)
            jobs = []
            try:
                for e in entries:
                    if e.data is None:
                        continue
                    lead = len(e.local_header(dos_time, dos_date)) + len(e.head)
                    src = np.frombuffer(e.data, dtype=np.uint8)
                    base = e.offset + lead
                    for lo in range(0, len(src), _PIECE):
                        hi = min(lo + _PIECE, len(src))
                        jobs.append(ex.submit(np.copyto, dst[base + lo:base + hi], src[lo:hi]))
                for j in jobs:
                    j.result()
                finish_crcs()
                for e in entries:
                    head = e.local_header(dos_time, dos_date) + e.head
                    mapped[e.offset:e.offset + len(head)] = head
                tail = directory()
                mapped[cd_offset:cd_offset + len(tail)] = tail
            finally:
                # every view of the mapping has to go before it can be closed
                for j in jobs:
                    j.cancel()
                for j in jobs:
                    if not j.cancelled():
                        j.exception()
                jobs.clear()
                src = dst = None
                try:
                    mapped.close()
                except BufferError:
                    # a propagating exception's traceback still holds a view of the mapping;
                    # the mapping goes with it
                    pass
        else:
            finish_crcs()
            with os.fdopen(os.dup(fd), "wb") as f:
                for e in entries:
                    f.write(e.local_header(dos_time, dos_date))
                    f.write(e.head)
                    if e.data is not None:
                        f.write(e.data)
                f.write(directory())
    finally:
        os.close(fd)
