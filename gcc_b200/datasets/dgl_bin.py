"""Reader / writer for DGL 0.4.x ``save_graphs`` files -- the on-disk input of the pretraining loader.

The reference keeps its pretraining corpus in ``data/small.bin`` written by
``dgl.data.utils.save_graphs(filename, g_list, labels={"graph_sizes": ...})`` (gcc/utils/x2dgl.py:129-131)
and reads it back with ``load_graphs(file, idx_list)`` per DataLoader worker and ``load_labels(file)`` in the
dataset constructor (gcc/datasets/graph_dataset.py:26-28,58-60).  DGL is not available here (and no real
``.bin`` file either), so the byte layout below is a restatement FROM MEMORY of DGL 0.4.3's
``src/graph/serialize/graph_serialize.cc`` / ``src/graph/immutable_graph.cc`` and dmlc-core's stream
conventions -- **parity unpinned until a file written by DGL itself is available**.  To be robust against the
details that cannot be checked, the reader only relies on

  * the file magic and the per-graph offset table, and
  * the NDArray record framing (magic, context, ndim, dtype, shape, byte size, payload),

and *locates* the three CSR arrays of every graph (indptr, indices, edge ids) by scanning the graph's byte range
for NDArray records, validating them structurally (indptr monotone from 0, last = len(indices), ids in range).
``write_dgl_bin`` produces the layout described below; ``tests/test_dgl_bin.py`` round-trips it and also feeds
the reader variants (extra header words, no graph-type word) to pin the tolerant behaviour.

Layout written (all little endian, dmlc ``Stream::Write`` conventions: vectors and strings are a uint64
count followed by the elements):

    uint64  0xDD2E4FF046B4A13F      kDGLSerializeMagic
    uint64  1                       version
    uint64  1                       graph type (kImmutableGraph)
    uint64  num_graph
    vector<uint64> graph_offsets    absolute file offset of every graph record
    vector<uint64> num_nodes, vector<uint64> num_edges
    vector<pair<string, NDArray>>   labels            ("graph_sizes" -> int64[num_graph])
    per graph:
        uint64  0xDD3C5FFE20046ABF  kDGLSerialize_ImGraph
        NDArray indptr (int64[n+1]), NDArray indices (int64[m]), NDArray edge_ids (int64[m])   -- the in-CSR
        vector<pair<string, NDArray>> node tensors, vector<pair<string, NDArray>> edge tensors (empty:
        x2dgl.py:121-123 clears ndata / edata before saving)

    NDArray record: uint64 0xDD5E40F096B4A13F, uint64 reserved, int32 device_type (1 = cpu), int32 device_id,
                    int32 ndim, uint8 dtype code (0 int, 1 uint, 2 float), uint8 bits, uint16 lanes,
                    int64 shape[ndim], int64 data_byte_size, payload
"""
import struct

import numpy as np

from . import synthetic

MAGIC_FILE = 0xDD2E4FF046B4A13F
MAGIC_IMGRAPH = 0xDD3C5FFE20046ABF
MAGIC_NDARRAY = 0xDD5E40F096B4A13F
_ND_MAGIC_BYTES = struct.pack("<Q", MAGIC_NDARRAY)
_DTYPES = {(0, 8): np.int8, (0, 16): np.int16, (0, 32): np.int32, (0, 64): np.int64,
           (1, 8): np.uint8, (1, 16): np.uint16, (1, 32): np.uint32, (1, 64): np.uint64,
           (2, 16): np.float16, (2, 32): np.float32, (2, 64): np.float64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class DglBinError(ValueError):
    pass


# ---- writer ------------------------------------------------------------------------------------------
def _pack_ndarray(a):
    a = np.ascontiguousarray(a)
    code, bits = _CODES[a.dtype]
    out = [struct.pack("<QQiii", MAGIC_NDARRAY, 0, 1, 0, a.ndim), struct.pack("<BBH", code, bits, 1)]
    out += [struct.pack("<q", s) for s in a.shape]
    out += [struct.pack("<q", a.nbytes), a.tobytes()]
    return b"".join(out)


def _pack_u64_vector(v):
    v = np.asarray(v, dtype=np.uint64)
    return struct.pack("<Q", len(v)) + v.tobytes()


def _pack_named_tensors(d):
    out = [struct.pack("<Q", len(d))]
    for name, arr in d.items():
        nb = name.encode()
        out += [struct.pack("<Q", len(nb)), nb, _pack_ndarray(np.asarray(arr))]
    return b"".join(out)


def write_dgl_bin(path, graphs, labels=None):
    """graphs: CSRGraph-likes (indptr, indices, num_nodes); labels: {name: ndarray}
    (default: {"graph_sizes": num_nodes per graph}, what x2dgl.py:129-131 stores)."""
    graphs = list(graphs)
    if labels is None:
        labels = {"graph_sizes": np.array([g.num_nodes for g in graphs], dtype=np.int64)}
    head = struct.pack("<QQQQ", MAGIC_FILE, 1, 1, len(graphs))
    nn = _pack_u64_vector([g.num_nodes for g in graphs])
    ne = _pack_u64_vector([len(g.indices) for g in graphs])
    lab = _pack_named_tensors(labels)
    blobs = []
    for g in graphs:
        indptr = np.asarray(g.indptr, dtype=np.int64)
        indices = np.asarray(g.indices, dtype=np.int64)
        eids = np.arange(len(indices), dtype=np.int64)
        blobs.append(struct.pack("<Q", MAGIC_IMGRAPH) + _pack_ndarray(indptr) + _pack_ndarray(indices) +
                     _pack_ndarray(eids) + _pack_named_tensors({}) + _pack_named_tensors({}))
    first = len(head) + 8 + 8 * len(graphs) + len(nn) + len(ne) + len(lab)
    offsets, pos = [], first
    for b in blobs:
        offsets.append(pos)
        pos += len(b)
    with open(path, "wb") as f:
        f.write(head + _pack_u64_vector(offsets) + nn + ne + lab + b"".join(blobs))
    return path


# ---- reader ------------------------------------------------------------------------------------------
def _parse_ndarray(buf, pos):
    """NDArray record at buf[pos:] -> (array, end) or None when the bytes there are not a plausible record."""
    if pos + 32 > len(buf) or buf[pos:pos + 8] != _ND_MAGIC_BYTES:
        return None
    dev_type, dev_id, ndim = struct.unpack_from("<iii", buf, pos + 16)
    if not (0 <= ndim <= 8) or dev_type not in (1, 2) or dev_id < 0:
        return None
    code, bits, lanes = struct.unpack_from("<BBH", buf, pos + 28)
    if (code, bits) not in _DTYPES or lanes != 1:
        return None
    p = pos + 32
    if p + 8 * (ndim + 1) > len(buf):
        return None
    shape = struct.unpack_from("<%dq" % ndim, buf, p) if ndim else ()
    p += 8 * ndim
    (nbytes,) = struct.unpack_from("<q", buf, p)
    p += 8
    count = int(np.prod(shape, dtype=np.int64)) if ndim else 1
    if any(s < 0 for s in shape) or nbytes != count * bits // 8 or p + nbytes > len(buf):
        return None
    arr = np.frombuffer(buf, dtype=_DTYPES[(code, bits)], count=count, offset=p).reshape(shape)
    return arr, p + nbytes


def _scan_ndarrays(buf, lo, hi, limit=None):
    """All NDArray records that start in buf[lo:hi], in file order (records never nest)."""
    out, pos = [], lo
    while True:
        pos = buf.find(_ND_MAGIC_BYTES, pos, hi)
        if pos < 0:
            break
        got = _parse_ndarray(buf, pos)
        if got is None:
            pos += 1
            continue
        out.append(got[0])
        pos = got[1]
        if limit and len(out) >= limit:
            break
    return out


def _csr_from_records(recs, n_hint=None):
    """First (indptr, indices) pair among the records that is structurally a CSR."""
    ints = [r for r in recs if r.ndim == 1 and r.dtype.kind in "iu"]
    for i in range(len(ints) - 1):
        indptr, indices = ints[i], ints[i + 1]
        if len(indptr) < 2 or indptr[0] != 0 or indptr[-1] != len(indices):
            continue
        if n_hint is not None and len(indptr) - 1 != n_hint:
            continue
        if np.any(np.diff(indptr) < 0) or (len(indices) and (indices.min() < 0 or indices.max() >= len(indptr) - 1)):
            continue
        return indptr.astype(np.int64), indices.astype(np.int64)
    raise DglBinError("no CSR (indptr, indices) pair found in a graph record")


def _read_header(buf):
    if len(buf) < 32 or struct.unpack_from("<Q", buf, 0)[0] != MAGIC_FILE:
        raise DglBinError("not a DGL save_graphs file (bad magic)")
    # after magic + version: an optional graph-type word, then num_graph followed by the offset vector whose
    # length word repeats num_graph -- accept either layout
    for skip in (3, 2, 4):
        p = 8 * skip
        if p + 16 > len(buf):
            continue
        num_graph, vec_len = struct.unpack_from("<QQ", buf, p)
        if 0 < num_graph == vec_len and p + 16 + 8 * num_graph <= len(buf):
            offsets = np.frombuffer(buf, dtype=np.uint64, count=int(num_graph), offset=p + 16).astype(np.int64)
            if np.all(offsets > p) and np.all(np.diff(offsets) > 0) and offsets[-1] < len(buf):
                return int(num_graph), offsets, p + 16 + 8 * int(num_graph)
    raise DglBinError("cannot locate the graph offset table")


def read_dgl_bin(path, idx_list=None):
    """-> (list of CSRGraph, labels dict).  idx_list: graph indices to load (load_graphs' second argument,
    graph_dataset.py:26-28); default all."""
    buf = open(path, "rb").read()
    num_graph, offsets, meta_end = _read_header(buf)
    ends = list(offsets[1:]) + [len(buf)]
    hints = None
    try:                                       # num_nodes vector follows the offset table
        (cnt,) = struct.unpack_from("<Q", buf, meta_end)
        if cnt == num_graph:
            hints = np.frombuffer(buf, dtype=np.uint64, count=num_graph, offset=meta_end + 8).astype(np.int64)
    except struct.error:
        pass
    labels = read_labels(path, _buf=buf)
    graphs = []
    for gi in (range(num_graph) if idx_list is None else idx_list):
        recs = _scan_ndarrays(buf, int(offsets[gi]), int(ends[gi]), limit=6)
        n_hint = int(hints[gi]) if hints is not None else None
        try:
            indptr, indices = _csr_from_records(recs, n_hint)
        except DglBinError:
            indptr, indices = _csr_from_records(recs, None)
        graphs.append(_as_out_csr(indptr, indices, "%s[%d]" % (path, gi)))
    return graphs, labels


def _as_out_csr(indptr, indices, name):
    """The record holds the IN-CSR (row = destination).  The reference's graphs are symmetric
    (x2dgl.py:40-62 adds both directions), where in-CSR == out-CSR; verify instead of assuming, and sort
    each neighbour list (the sampler's contract, gccb200.h gccb_graph_t)."""
    n = len(indptr) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    key = rows * n + indices
    order = np.argsort(key, kind="stable")
    key_sorted = key[order]
    rev = np.sort(indices * n + rows)
    if not np.array_equal(key_sorted, rev):
        raise DglBinError("%s: graph is not symmetric -- the pretraining path expects the output of x2dgl.py "
                          "(bidirected, gcc/utils/x2dgl.py:40-62)" % name)
    return synthetic.CSRGraph(indptr.copy(), indices[order].astype(np.int32), n, name)


def read_labels(path, _buf=None):
    """load_labels(file) (graph_dataset.py:58-60): {name: ndarray}; the reference reads "graph_sizes"."""
    buf = _buf if _buf is not None else open(path, "rb").read()
    num_graph, offsets, meta_end = _read_header(buf)
    lo, hi = meta_end, int(offsets[0])
    labels, pos = {}, lo
    while True:                                 # a label = string (uint64 length + bytes) right before an NDArray
        pos = buf.find(_ND_MAGIC_BYTES, pos, hi)
        if pos < 0:
            break
        got = _parse_ndarray(buf, pos)
        if got is None:
            pos += 1
            continue
        name = None
        for ln in range(1, 65):
            s = pos - ln
            if s - 8 < lo:
                break
            if struct.unpack_from("<Q", buf, s - 8)[0] == ln:
                cand = buf[s:pos]
                if all(32 <= c < 127 for c in cand):
                    name = cand.decode()
                    break
        labels[name or "label%d" % len(labels)] = got[0].copy()
        pos = got[1]
    return labels
