"""
Host-side input readers of the drop-in class (SURVEY.md section 8(f) row 1:
with the arithmetic on the GPU, reading the input files is the critical path
of a files-in -> files-out run).

``load_npz`` reads what ``scipy.sparse.save_npz`` writes -- the reference's
input format, ``hic3defdr/analysis/analysis.py:85-86`` loads it with
``scipy.sparse.load_npz`` -- without numpy's chunked zip reader: every zip
member is inflated by ONE zlib call straight into its final buffer (zlib drops
the GIL for the whole member, so the members of a file and the files of
several chromosomes inflate in parallel on a thread pool), and the arrays are
views of those buffers.  Anything this reader does not recognise goes to
``scipy.sparse.load_npz``.
"""
import io
import struct
import zipfile
import zlib

import numpy as np
import scipy.sparse as sparse
from numpy.lib import format as npformat

_LOCAL_HEADER = struct.Struct('<4s5H3L2H')     # zip local file header, 30 bytes
_MEMBERS = {'csr': ('data', 'indices', 'indptr'),
            'csc': ('data', 'indices', 'indptr'),
            'bsr': ('data', 'indices', 'indptr'),
            'coo': ('data', 'row', 'col'),
            'dia': ('data', 'offsets')}


class Unsupported(Exception):
    pass


def _member_bytes(handle, info):
    """compressed bytes of one zip member."""
    handle.seek(info.header_offset)
    head = _LOCAL_HEADER.unpack(handle.read(_LOCAL_HEADER.size))
    if head[0] != b'PK\x03\x04':
        raise Unsupported('bad local header')
    handle.seek(info.header_offset + _LOCAL_HEADER.size + head[9] + head[10])
    return handle.read(info.compress_size)


def _inflate(raw, info):
    if info.compress_type == zipfile.ZIP_STORED:
        buf = raw
    elif info.compress_type == zipfile.ZIP_DEFLATED:
        buf = zlib.decompress(raw, -15, max(info.file_size, 1))
    else:
        raise Unsupported('compression method %d' % info.compress_type)
    if len(buf) != info.file_size or \
            (zlib.crc32(buf) & 0xffffffff) != info.CRC:
        raise ValueError('corrupt member %s' % info.filename)
    return buf


def _array(buf):
    """the ndarray stored in the bytes of one ``.npy`` member (a read-only
    view of ``buf`` for plain dtypes)."""
    head = io.BytesIO(buf[:4096] if len(buf) > 4096 else buf)
    version = npformat.read_magic(head)
    if version == (1, 0):
        shape, fortran, dtype = npformat.read_array_header_1_0(head)
    elif version == (2, 0):
        shape, fortran, dtype = npformat.read_array_header_2_0(head)
    else:
        raise Unsupported('npy version %r' % (version,))
    if dtype.hasobject:
        raise Unsupported('object array')
    count = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    a = np.frombuffer(buf, dtype=dtype, count=count, offset=head.tell())
    return a.reshape(shape[::-1]).T if fortran else a.reshape(shape)


def _read_members(path, names, pool):
    with open(path, 'rb') as handle:
        with zipfile.ZipFile(handle) as z:
            infos = {i.filename: i for i in z.infolist()}
        if '_is_array.npy' in infos:      # scipy >= 1.11 sparse arrays
            raise Unsupported('sparse array container')
        want = []
        for n in names:
            info = infos.get(n + '.npy')
            if info is None:
                raise Unsupported('no member %s' % n)
            if info.flag_bits & 0x1:
                raise Unsupported('encrypted')
            want.append((info, _member_bytes(handle, info)))
    if pool is None:
        return [_array(_inflate(raw, info)) for info, raw in want]
    futures = [pool.submit(_inflate, raw, info) for info, raw in want]
    return [_array(f.result()) for f in futures]


def load_npz(path, pool=None):
    """``scipy.sparse.load_npz(path)``; ``pool`` (a ThreadPoolExecutor the
    caller is NOT itself running on) inflates the members in parallel."""
    try:
        fmt, shape = _read_members(path, ('format', 'shape'), None)
        fmt = fmt.item()
        if isinstance(fmt, bytes):
            fmt = fmt.decode('ascii')
        names = _MEMBERS.get(fmt)
        if names is None:
            raise Unsupported('format %r' % (fmt,))
        parts = _read_members(path, names, pool)
    except (Unsupported, zipfile.BadZipFile, struct.error):
        return sparse.load_npz(path)
    shape = tuple(int(v) for v in shape)
    cls = getattr(sparse, fmt + '_matrix')
    if fmt == 'coo':
        data, row, col = parts
        return cls((data, (row, col)), shape=shape)
    if fmt == 'dia':
        data, offsets = parts
        return cls((data, offsets), shape=shape)
    data, indices, indptr = parts
    return cls((data, indices, indptr), shape=shape)


def pin_csr(m, device=None):
    """The CSR matrix ``m`` re-based on page-locked host memory: the same
    matrix (scipy object over numpy views of pinned torch tensors, canonical
    flag carried over) whose arrays ``ops.DeviceCSR`` uploads with asynchronous
    copies instead of blocking pageable ones.  The pinned tensors ride along
    as ``m._h3d_pinned`` (data, indices, indptr).  Returns ``m`` itself when
    there is no CUDA device or a dtype torch cannot hold."""
    import torch
    if not torch.cuda.is_available():
        return m
    if device is not None:
        torch.cuda.set_device(device)      # loader threads start on device 0
    parts = []
    try:
        for a in (m.data, m.indices, m.indptr):
            a = np.ascontiguousarray(a)
            dtype = torch.from_numpy(np.empty(0, dtype=a.dtype)).dtype
            t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
            np.copyto(t.numpy(), a)
            parts.append(t)
    except (TypeError, RuntimeError):
        return m
    out = sparse.csr_matrix(tuple(t.numpy() for t in parts), shape=m.shape)
    if m.has_canonical_format:
        out.has_canonical_format = True
    out._h3d_pinned = tuple(parts)
    return out
