"""Graph containers and loaders for the hot path.

PyG (``torch_geometric.data.Data`` / ``Batch`` / ``NeighborLoader`` /
``DataLoader``) is what the reference's ``fit`` feeds the models
(pygda/models/a2gnn.py:254-288); it is not installable here, so these are the
duck-typed equivalents the estimators in ``pygda_b200.models`` accept.  Any
object with ``.x [N,F]``, ``.edge_index [2,E] int64``, ``.y`` (and ``.batch``
in graph mode) and ``.to(device)`` works, including real PyG objects.
"""
import ctypes as C

import numpy as np
import torch


class PackedRows:
    """Row-compressed PINNED host copy of a sparse fp32 matrix (bag-of-words features are a few per cent
    non-zero): the non-zero values and their DELTA-coded column ids, one byte per entry (within a row the running
    column starts at 0, every byte adds its value, 255 is an escape that only advances) -- 5 bytes per non-zero
    instead of the dense row.  ``to_dense(device)`` copies only these over PCIe and rebuilds the dense [N, F] matrix
    on the GPU bit for bit (``gda_unpack_rows_delta_f32``; an entry is kept iff its BIT PATTERN is non-zero, so
    -0.0 survives)."""

    def __init__(self, x, chunk=8192):
        n, f = x.shape
        x = x.contiguous()
        vals, deltas, counts, nbytes = [], [], [], []
        for s in range(0, n, chunk):
            blk = x[s:s + chunk]
            mask = blk.view(torch.int32) != 0
            cnt = mask.sum(1)
            col = mask.nonzero(as_tuple=True)[1]
            vals.append(blk[mask])
            counts.append(cnt)
            # gap to the previous entry of the same row (to column 0 for a row's first entry)
            d = col.clone()
            if col.numel() > 1:
                d[1:] -= col[:-1]
            starts = (torch.cumsum(cnt, 0) - cnt)[cnt > 0]
            d[starts] = col[starts]
            width = 1 + d // 255                                  # escape bytes + the emitting byte
            pos = torch.cumsum(width, 0) - 1
            out = torch.full((int(width.sum()),), 255, dtype=torch.uint8)
            out[pos] = (d % 255).to(torch.uint8)
            deltas.append(out)
            rows = torch.repeat_interleave(torch.arange(blk.size(0)), cnt)
            nbytes.append(torch.zeros(blk.size(0), dtype=torch.int64).index_add_(0, rows, width))
        def ptr(parts):
            p = torch.zeros(n + 1, dtype=torch.int64)
            if parts:
                p[1:] = torch.cumsum(torch.cat(parts), 0)
            return p
        val_ptr, byte_ptr = ptr(counts), ptr(nbytes)
        if int(byte_ptr[-1]) >= 2 ** 31:
            raise ValueError("PackedRows: more than 2^31 packed bytes; pin_memory(pack=False) keeps the dense form")
        self.shape = (n, f)
        pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
        self.vals = pin(torch.cat(vals) if vals else torch.zeros(0))
        self.deltas = pin(torch.cat(deltas) if deltas else torch.zeros(0, dtype=torch.uint8))
        self.val_ptr = pin(val_ptr.to(torch.int32))
        self.byte_ptr = pin(byte_ptr.to(torch.int32))

    def tensors(self):
        """The pinned host arrays that cross PCIe, by staging name."""
        return {"_vals": self.vals, "_deltas": self.deltas, "_val_ptr": self.val_ptr, "_byte_ptr": self.byte_ptr}

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors().values())

    def unpack_into(self, dev, out, stream=None):
        """Rebuild the dense matrix in ``out`` [N, F] from DEVICE copies ``dev`` of ``tensors()`` (current stream)."""
        from ._lib import gda
        n, f = self.shape
        p = lambda t: C.c_void_p(t.data_ptr())                    # noqa: E731
        gda.unpack_rows_delta_f32(p(dev["_vals"]), p(dev["_deltas"]), p(dev["_val_ptr"]), p(dev["_byte_ptr"]), n, f,
                                  p(out), f, C.c_void_p(0),
                                  C.c_void_p((stream or torch.cuda.current_stream(out.device)).cuda_stream))

    def to_dense(self, device, non_blocking=True):
        n, f = self.shape
        dev = torch.device(device)
        d = {k: t.to(dev, non_blocking=non_blocking) for k, t in self.tensors().items()}
        out = torch.empty(n, f, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            self.unpack_into(d, out)
        for t in d.values():
            t.record_stream(torch.cuda.current_stream(dev))
        return out


class XTiles:
    """DEVICE view of a tile-packed sparse matrix (``PackedTiles``): what ``gda_gemm_xt_fwd`` / ``gda_gemm_xt_dw``
    read.  Attached to a dense device ``x`` as ``x._gda_tiles`` (``Data.to``) so that the first layer's products
    run from the packed form (ops.GraphConvFn / ops.LinearFn)."""
    __slots__ = ("vals", "codes", "ptr", "seg", "rows", "cols", "dense_version")

    def __init__(self, vals, codes, ptr, seg, rows, cols):
        self.vals, self.codes, self.ptr, self.seg = vals, codes, ptr, seg
        self.rows, self.cols = int(rows), int(cols)
        # ``_version`` of the dense tensor this view is attached to, at attach time: an in-place edit of the dense
        # matrix afterwards makes the packed copy stale, and ops.x_tiles then ignores it
        self.dense_version = None

    def tensors(self):
        return {"_vals": self.vals, "_codes": self.codes, "_ptr": self.ptr, "_seg": self.seg}

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors().values())

    def unpack_into(self, out, stream=None):
        """Rebuild the dense [rows, cols] matrix in ``out`` (current stream): ``gda_unpack_tiles_f32``."""
        from ._lib import gda
        p = lambda t: C.c_void_p(t.data_ptr())                    # noqa: E731
        gda.unpack_tiles_f32(p(self.vals), p(self.codes), p(self.ptr), p(self.seg), self.rows, self.cols,
                             p(out), out.stride(0), C.c_void_p(0),
                             C.c_void_p((stream or torch.cuda.current_stream(out.device)).cuda_stream))
        return out


class PackedTiles:
    """TILE-PACKED copy of a sparse fp32 matrix -- the operand form of the first layer's tensor-core GEMMs
    (csrc/gemm_xt.cu) AND the pinned staging form of the per-step ``.to(device)`` (pygda/models/a2gnn.py:311-312):
    the host->device copy lands in the buffers the kernels read; the dense matrix is not needed on the path.

    Sub-tile ``t = (row // 32) * nkb + col // 64`` (``nkb = ceil(cols / 64)``); inside a sub-tile the entries are
    sorted by ``p = (row % 32) * 64 + col % 64``.  Entries of sub-tile ``t``: ``[ptr[t], ptr[t + 1])`` of ``vals``
    (fp32) and ``codes`` (uint8 ``= p & 255``); ``seg[t, k]`` (int16 bit pattern of a uint16, ``k = 0..7``) = number
    of entries of the sub-tile with ``p < 256 (k + 1)``, so entry ``i`` has ``p >> 8 = #{k < 7: seg[t, k] <= i}``:
    5 bytes per non-zero + 20 bytes per sub-tile, and every entry decodes on its own (no prefix sums in the
    kernels).  An entry is kept iff its BIT PATTERN is non-zero (-0.0 survives).  The strip count is padded to a
    multiple of 4 (one 128-row operand tile).  Built with torch ops on whatever device ``x`` lives on (one-time
    preparation, like the CSR of the graph)."""

    def __init__(self, x, chunk=8192, pin=True, compress_values=None):
        n, f = x.shape
        x = x.contiguous()
        dev = x.device
        # the pinned HOST staging form also packs the exponents of the values (3.5 instead of 4 bytes per value,
        # lossless: ``compress``); a device-built copy is the operand form itself and keeps plain fp32 values
        self.compressed = (pin and dev.type == "cpu") if compress_values is None else bool(compress_values)
        nkb = -(-f // 64)
        nstrips = -(-(-(-n // 32)) // 4) * 4
        ntiles = nstrips * nkb
        chunk = max(32, chunk // 32 * 32)
        vals, codes, cnt8 = [], [], []
        for s in range(0, n, chunk):
            blk = x[s:s + chunk]
            mask = blk.view(torch.int32) != 0
            r, c = mask.nonzero(as_tuple=True)
            key = ((r // 32) * nkb + c // 64) * 2048 + (r % 32) * 64 + c % 64
            order = torch.argsort(key)
            key = key[order]
            vals.append(blk[mask][order])
            codes.append((key & 255).to(torch.uint8))
            ntl = -(-blk.size(0) // 32) * nkb
            cnt8.append(torch.bincount(key >> 8, minlength=ntl * 8))     # per (sub-tile, 256-position segment)
        cnt = torch.zeros(ntiles, 8, dtype=torch.int64, device=dev)
        if cnt8:
            flat = torch.cat(cnt8)
            cnt.view(-1)[:flat.numel()] = flat
        seg = torch.cumsum(cnt, 1)
        ptr = torch.zeros(ntiles + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(seg[:, 7], 0)
        if int(ptr[-1]) >= 2 ** 31:
            raise ValueError("PackedTiles: more than 2^31 non-zeros; pin_memory(pack=False) keeps the dense form")
        self.shape = (n, f)
        pin_ = (lambda t: t.pin_memory()) if (pin and torch.cuda.is_available() and dev.type == "cpu") else (lambda t: t)
        self._pin = pin_
        self.vals = torch.cat(vals) if vals else torch.zeros(0, device=dev)
        if not self.compressed:
            self.vals = pin_(self.vals)
        self.codes = pin_(torch.cat(codes) if codes else torch.zeros(0, dtype=torch.uint8, device=dev))
        self.ptr = pin_(ptr.to(torch.int32))
        self.seg = pin_(seg.to(torch.int16).contiguous())           # counts <= 2048: the int16 bits ARE the uint16
        if self.compressed:
            self.compress()

    def compress(self):
        """(Re)build the exponent-packed staging arrays from ``self.vals``: per value 3 bytes ``sign << 23 | mantissa``
        (``m24``) and a 4-bit code ``exponent - base`` (``ecode``, 15 = escape, value kept verbatim in ``esc_idx`` /
        ``esc_val``); ``vmeta = [base, number of escapes, 0, 0]``.  ``base`` is the 15-binade window that covers most
        values.  Bit-exact for any fp32 input (``gda_unpack_values_f32`` / ``decompress_values``)."""
        v = self.vals
        nnz = v.numel()
        g4 = -(-nnz // 4)
        bits = v.view(torch.int32)
        exp = (bits >> 23) & 0xFF
        v24 = (bits & 0x7FFFFF) | (((bits >> 31) & 1) << 23)
        hist = torch.bincount(exp, minlength=256).to(torch.int64)
        win = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int64), hist]), 0)
        cover = win[15:256 + 1] - win[0:256 - 15 + 1]               # values with exponent in [base, base + 15)
        base = int(torch.argmax(cover)) if nnz else 0
        code = exp - base
        esc = (code < 0) | (code > 14)
        code = torch.where(esc, torch.full_like(code, 15), code)
        m24 = torch.zeros(4 * g4, 3, dtype=torch.uint8)
        m24[:nnz, 0] = (v24 & 255).to(torch.uint8)
        m24[:nnz, 1] = ((v24 >> 8) & 255).to(torch.uint8)
        m24[:nnz, 2] = ((v24 >> 16) & 255).to(torch.uint8)
        c4 = torch.zeros(4 * g4, dtype=torch.int32)
        c4[:nnz] = code
        ecode = (c4[0::2] | (c4[1::2] << 4)).to(torch.uint8)
        esc_idx = esc.nonzero(as_tuple=True)[0].to(torch.int32)
        n_esc = int(esc_idx.numel())
        cap = max(n_esc, 1)
        new = {"m24": m24.reshape(-1), "ecode": ecode.contiguous(),
               "vmeta": torch.tensor([base, n_esc, 0, 0], dtype=torch.int32),
               "esc_idx": torch.zeros(cap, dtype=torch.int32), "esc_val": torch.zeros(cap, dtype=torch.float32)}
        new["esc_idx"][:n_esc] = esc_idx
        new["esc_val"][:n_esc] = v[esc]
        for k, t in new.items():
            old = getattr(self, k, None)
            if old is not None and old.shape == t.shape:
                old.copy_(t)                                         # keep the pinned buffers a staging plan holds
            else:
                setattr(self, k, self._pin(t))

    def decompress_values(self):
        """fp32 values rebuilt from the exponent-packed arrays on the host (tests; the device does it in
        ``gda_unpack_values_f32``)."""
        nnz = self.vals.numel()
        m = self.m24.view(-1, 3)[:nnz].to(torch.int32)
        v24 = m[:, 0] | (m[:, 1] << 8) | (m[:, 2] << 16)
        e = self.ecode.to(torch.int32)
        code = torch.stack([e & 15, e >> 4], 1).reshape(-1)[:nnz]
        base, n_esc = int(self.vmeta[0]), int(self.vmeta[1])
        bits = ((v24 >> 23) << 31) | ((base + code) << 23) | (v24 & 0x7FFFFF)
        out = bits.view(torch.float32).clone()
        out[self.esc_idx[:n_esc].long()] = self.esc_val[:n_esc]
        return out

    def tensors(self):
        """The arrays that cross PCIe, by staging name."""
        t = {"_codes": self.codes, "_ptr": self.ptr, "_seg": self.seg}
        if self.compressed:
            t.update({"_m24": self.m24, "_ecode": self.ecode, "_vmeta": self.vmeta, "_esc_idx": self.esc_idx,
                      "_esc_val": self.esc_val})
        else:
            t["_vals"] = self.vals
        return t

    def unpack_values_into(self, dev, vals, stream=None):
        """fp32 ``vals`` (device, [nnz]) from DEVICE copies ``dev`` of the exponent-packed arrays (current stream)."""
        from ._lib import gda
        p = lambda t: C.c_void_p(t.data_ptr())                    # noqa: E731
        gda.unpack_values_f32(p(dev["_m24"]), p(dev["_ecode"]), p(dev["_vmeta"]), p(dev["_esc_idx"]),
                              p(dev["_esc_val"]), dev["_esc_idx"].numel(), vals.numel(), p(vals),
                              C.c_void_p((stream or torch.cuda.current_stream(vals.device)).cuda_stream))
        return vals

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors().values())

    def view(self, dev=None):
        """``XTiles`` over ``dev`` (DEVICE copies of ``tensors()``; default: this object's own arrays)."""
        d = dev if dev is not None else self.tensors()
        if "_vals" in d:
            vals = d["_vals"]
        else:                                                   # exponent-packed staging copy: rebuild the fp32 values
            vals = torch.empty(self.vals.numel(), dtype=torch.float32, device=d["_codes"].device)
            self.unpack_values_into(d, vals)
        return XTiles(vals, d["_codes"], d["_ptr"], d["_seg"], *self.shape)

    def unpack_into(self, dev, out, stream=None):
        """Rebuild the dense matrix in ``out`` [N, F] from DEVICE copies ``dev`` of ``tensors()`` (current stream)."""
        self.view(dev).unpack_into(out, stream)

    @property
    def staged_nbytes(self):
        return self.nbytes

    def to_dense(self, device, non_blocking=True):
        """Dense device matrix rebuilt from a fresh host->device copy of the packed arrays; the copy stays attached
        as ``._gda_tiles`` -- the first layer multiplies from it, the dense form serves every other consumer."""
        n, f = self.shape
        dev = torch.device(device)
        d = {k: t.to(dev, non_blocking=non_blocking) for k, t in self.tensors().items()}
        out = torch.empty(n, f, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            tiles = self.view(d)
            tiles.unpack_into(out)
        for t in list(d.values()) + [tiles.vals]:
            t.record_stream(torch.cuda.current_stream(dev))
        out._gda_tiles = tiles
        tiles.dense_version = out._version
        return out

    def decode(self):
        """Dense matrix by a plain sequential decode on the host (tests; not a product path)."""
        n, f = self.shape
        nkb = -(-f // 64)
        out = np.zeros((n, f), dtype=np.float32)
        vals, codes = self.vals.cpu().numpy(), self.codes.cpu().numpy()
        ptr, seg = self.ptr.cpu().numpy(), self.seg.cpu().numpy().astype(np.int64)
        for t in range(len(ptr) - 1):
            for i in range(ptr[t + 1] - ptr[t]):
                pos = int((seg[t, :7] <= i).sum()) * 256 + int(codes[ptr[t] + i])
                out[(t // nkb) * 32 + pos // 64, (t % nkb) * 64 + pos % 64] = vals[ptr[t] + i]
        return torch.from_numpy(out)


PACK_DENSITY = 0.25      # pin_memory() keeps x row-compressed below this fraction of non-zeros


class Data:
    def __init__(self, x=None, edge_index=None, y=None, batch=None, num_graphs=None, **kw):
        self.x = x
        self.edge_index = edge_index
        self.y = y
        self.batch = batch
        self.num_graphs = num_graphs
        for k, v in kw.items():
            setattr(self, k, v)

    # ---- PyG-compatible surface used by the reference's estimators ----
    def to(self, device, non_blocking=False):
        dev = torch.device(device)
        if all((not torch.is_tensor(v)) or v.device == dev for v in self.__dict__.values()):
            return self                      # same no-op PyG performs when already resident
        out = self.__class__.__new__(self.__class__)
        packed = self.__dict__.get("_packed_x") if dev.type == "cuda" else None
        for k, v in self.__dict__.items():
            if k == "_packed_x":
                continue
            if k == "x" and packed is not None:
                out.__dict__[k] = packed.to_dense(dev)          # only the non-zeros cross PCIe
                continue
            out.__dict__[k] = v.to(dev, non_blocking=non_blocking) if torch.is_tensor(v) else v
        ei, src = out.__dict__.get("edge_index"), self.edge_index
        if torch.is_tensor(ei) and ei is not src:
            # The normalised CSR is a pure function of edge_index: key the device copy by the
            # tensor it was copied from, so that re-sending the same host graph every step
            # (what the reference's fit loop does, a2gnn.py:311-312) maps to one gda_graph_t.
            ei._gda_key = getattr(src, "_gda_key", None) or (
                "src", src.data_ptr(), src._version, tuple(src.shape), str(src.device))
            ei._gda_keepalive = getattr(src, "_gda_keepalive", src)
            part = getattr(src, "_gda_partition", None)      # row-partitioned multi-GPU graph tag
            if part is not None:
                ei._gda_partition = part
        xd, xsrc = out.__dict__.get("x"), self.__dict__.get("x")
        if torch.is_tensor(xd) and xd is not xsrc and torch.is_tensor(xsrc) and dev.type == "cuda":
            # same for the input features: their derived operand forms (split-bf16 pair / bf16 copy, ops.ConstCache)
            # are keyed by the host tensor, so re-sending it every step refills ONE set of buffers
            xd._gda_key = getattr(xsrc, "_gda_key", None) or (
                "src", xsrc.data_ptr(), xsrc._version, tuple(xsrc.shape), str(xsrc.device))
            xd._gda_keepalive = getattr(xsrc, "_gda_keepalive", xsrc)
        yd, ysrc = out.__dict__.get("y"), self.__dict__.get("y")
        if torch.is_tensor(yd) and yd is not ysrc and torch.is_tensor(ysrc):
            yd._gda_src = ysrc                      # label-range validation is remembered on the host tensor (ops)
        ew, wsrc = out.__dict__.get("edge_weight"), self.__dict__.get("edge_weight")
        if torch.is_tensor(ew) and ew is not wsrc:
            # same for per-edge weights (StruRW): the re-weighted CSR is keyed by the host tensor they came from
            ew._gda_key = getattr(wsrc, "_gda_key", None) or ("src", wsrc.data_ptr(), wsrc._version, str(wsrc.device))
            ew._gda_keepalive = getattr(wsrc, "_gda_keepalive", wsrc)
        return out

    def pin_memory(self, pack="auto"):
        """Pinned host copy for the per-step ``.to(device)`` of the fit loops.  ``pack``: keep a sparse fp32 ``x``
        compressed in the pinned staging area -- "auto" does so below ``PACK_DENSITY`` non-zeros, tile-packed
        (``PackedTiles``: the form the first layer's GEMMs read; "rows" selects the older row-compressed
        ``PackedRows``); ``.x`` of the result stays the caller's host tensor, ``.to(device)`` rebuilds it densely
        and keeps the packed device copy attached for the first layer."""
        out = self.__class__.__new__(self.__class__)
        x = self.__dict__.get("x")
        do_pack = False
        if pack and torch.is_tensor(x) and not x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.numel():
            if pack == "auto":
                probe = x[:: max(1, x.size(0) // 512)]
                do_pack = float((probe.reshape(-1).view(torch.int32) != 0).float().mean()) < PACK_DENSITY
            else:
                do_pack = True
        for k, v in self.__dict__.items():
            if k == "x" and do_pack:
                out.__dict__[k] = v
                continue
            pinnable = torch.is_tensor(v) and not v.is_cuda and torch.cuda.is_available()
            out.__dict__[k] = v.pin_memory() if pinnable else v
        if do_pack:
            out.__dict__["_packed_x"] = PackedRows(x) if pack == "rows" else PackedTiles(x)
        src, ei = self.edge_index, out.__dict__.get("edge_index")
        if torch.is_tensor(ei) and ei is not src:
            for attr in ("_gda_partition", "_gda_key", "_gda_keepalive"):
                if hasattr(src, attr):
                    setattr(ei, attr, getattr(src, attr))
        return out

    def h2d_nbytes(self):
        """Bytes ``.to(cuda)`` moves over PCIe for this (host) object."""
        total = 0
        for k, v in self.__dict__.items():
            if k == "x" and "_packed_x" in self.__dict__:
                total += self.__dict__["_packed_x"].nbytes
            elif torch.is_tensor(v) and not v.is_cuda:
                total += v.numel() * v.element_size()
        return total

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_features(self):
        return self.x.size(1)

    def __len__(self):
        # ``len(Batch)`` is the number of graphs (used by pygda/models/a2gnn.py:270-271)
        return self.num_graphs if self.num_graphs is not None else 1

    def __repr__(self):
        parts = [f"{k}={list(v.shape)}" for k, v in self.__dict__.items() if torch.is_tensor(v)]
        return "Data(" + ", ".join(parts) + ")"


class Batch(Data):
    @staticmethod
    def from_data_list(graphs):
        xs, eis, ys, bs, ptr = [], [], [], [], [0]
        off = 0
        for g, d in enumerate(graphs):
            xs.append(d.x)
            eis.append(d.edge_index + off)
            ys.append(d.y.view(-1))
            bs.append(torch.full((d.x.size(0),), g, dtype=torch.long, device=d.x.device))
            off += d.x.size(0)
            ptr.append(off)
        return Batch(x=torch.cat(xs), edge_index=torch.cat(eis, 1), y=torch.cat(ys),
                     batch=torch.cat(bs), num_graphs=len(graphs),
                     ptr=torch.tensor(ptr, dtype=torch.long))


class NeighborLoader:
    """Full-batch stand-in for ``NeighborLoader(data, [-1]*L, batch_size=N)``:
    one batch per epoch = the whole graph, node order preserved, edges regrouped
    by destination (stable), exactly what PyG yields when every node is a seed
    (SURVEY.md Appendix A.6).  Sampled mini-batching (``batch_size < N``) is the
    reference's own scaling tool and is replaced here by node partitioning
    (DESIGN.md section 6); asking for it raises."""

    def __init__(self, data, num_neighbors, batch_size=None, pin=False, prefetch_device=None, **kw):
        n = data.x.shape[0]
        if batch_size is not None and batch_size != n:
            raise NotImplementedError(
                "pygda_b200 runs node-level models full-batch (batch_size=0); "
                "neighbour-sampled mini-batches are not part of the accelerated path")
        self.data = data
        if torch.is_tensor(data.x) and data.x.is_cuda:
            data.x._gda_const = True            # resident input features: operand forms are cached (ops.ConstCache)
        ei = data.edge_index
        order = torch.argsort(ei[1], stable=True)
        extra = {k: v for k, v in data.__dict__.items()
                 if k not in ("x", "edge_index", "y", "batch", "num_graphs")}
        extra.pop("_packed_x", None)
        for k in ("edge_weight", "edge_attr"):      # edge-level attributes travel with their edges [upstream filter_data]
            v = extra.get(k)
            if torch.is_tensor(v) and v.dim() >= 1 and v.size(0) == ei.size(1):
                extra[k] = v[order.to(v.device)].contiguous()
        self._batch = Data(x=data.x, edge_index=ei[:, order].contiguous(), y=data.y,
                           batch=data.batch, num_graphs=data.num_graphs, **extra)
        if hasattr(ei, "_gda_partition"):            # row-partitioned multi-GPU graph: the tag travels with the edges
            self._batch.edge_index._gda_partition = ei._gda_partition
        # the fit loops send this one batch host->device on EVERY step (pygda/models/a2gnn.py:311-312): stage it in
        # pinned memory once (a sparse x row-compressed, Data.pin_memory) instead of paging it through each time
        host = torch.is_tensor(data.x) and not data.x.is_cuda
        if pin and torch.cuda.is_available() and host:
            self._batch = self._batch.pin_memory()
        # Opt-in software pipeline for host-resident graphs: the copy of the NEXT epoch's batch is issued on a side
        # stream when the current one is handed out, so PCIe traffic overlaps the training step instead of preceding
        # it (the reference's loop copies, then computes: a2gnn.py:311-319).  Same bytes, same values.
        self._prefetch_device = None
        if prefetch_device is not None and pin and host and torch.cuda.is_available() \
                and torch.device(prefetch_device).type == "cuda":
            self._prefetch_device = torch.device(prefetch_device)
        self._staged, self._side = None, None

    def _stage(self):
        dev = self._prefetch_device
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
        with torch.cuda.stream(self._side):
            d = self._batch.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._side)
        return d, ev

    def _iter_prefetched(self):
        dev = self._prefetch_device
        if self._staged is None:
            self._staged = self._stage()
        d, ev = self._staged
        self._staged = self._stage()                     # next epoch's copy starts now
        main = torch.cuda.current_stream(dev)
        main.wait_event(ev)
        for v in d.__dict__.values():                    # allocated on the side stream, consumed on this one
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(main)
                tiles = getattr(v, "_gda_tiles", None)   # the packed copy the first layer reads travels with x
                if tiles is not None:
                    for t in tiles.tensors().values():
                        t.record_stream(main)
        yield d

    def __iter__(self):
        if self._prefetch_device is not None:
            yield from self._iter_prefetched()
            return
        # PyG builds a NEW Data object for every batch: attributes a fit loop sets on the batch it was handed
        # (StruRW's re-weighted ``edge_weight``, pygda/models/strurw.py:487) do not survive into the next epoch.
        # The copy is shallow -- same tensors, so graph / split caches keyed by tensor identity keep hitting.
        out = self._batch.__class__.__new__(self._batch.__class__)
        out.__dict__.update(self._batch.__dict__)
        yield out

    def __len__(self):
        return 1


class DeviceGraphDataset:
    """A list of small graphs kept RESIDENT on the GPU as one concatenation (x_all, edge_index_all with global node
    ids, node_ptr / edge_ptr), so that a mini-batch is collated by two kernels (``gda_collate_graphs``) instead
    of a Python loop + ``torch.cat`` over hundreds of graphs per step (``Batch.from_data_list``).  The result is
    the same ``Batch`` (tests/test_gpu_loader.py)."""

    def __init__(self, graphs, device):
        self.device = torch.device(device)
        graphs = list(graphs)
        self.num_graphs = len(graphs)
        nn_ = torch.tensor([g.x.size(0) for g in graphs], dtype=torch.long)
        ne_ = torch.tensor([g.edge_index.size(1) for g in graphs], dtype=torch.long)
        self.node_ptr_host = torch.zeros(self.num_graphs + 1, dtype=torch.long)
        self.edge_ptr_host = torch.zeros(self.num_graphs + 1, dtype=torch.long)
        self.node_ptr_host[1:] = torch.cumsum(nn_, 0)
        self.edge_ptr_host[1:] = torch.cumsum(ne_, 0)
        offs = self.node_ptr_host[:-1].tolist()
        self.x_all = torch.cat([g.x for g in graphs]).to(self.device, torch.float32).contiguous()
        self.ei_all = torch.cat([g.edge_index.cpu() + o for g, o in zip(graphs, offs)], 1).to(self.device).contiguous()
        self.y_all = torch.cat([g.y.view(-1) for g in graphs]).to(self.device)
        self.node_ptr, self.edge_ptr = self.node_ptr_host.to(self.device), self.edge_ptr_host.to(self.device)
        self.num_features = self.x_all.size(1)

    def __len__(self):
        return self.num_graphs

    def collate(self, ids):
        """``Batch`` of the graphs ``ids`` (a list / CPU tensor, in batch order), built on the device."""
        import ctypes as C
        from ._lib import gda
        ids_h = torch.as_tensor(ids, dtype=torch.long)
        b = ids_h.numel()
        out_np = torch.zeros(b + 1, dtype=torch.long)
        out_ep = torch.zeros(b + 1, dtype=torch.long)
        out_np[1:] = torch.cumsum(self.node_ptr_host[ids_h + 1] - self.node_ptr_host[ids_h], 0)
        out_ep[1:] = torch.cumsum(self.edge_ptr_host[ids_h + 1] - self.edge_ptr_host[ids_h], 0)
        n_out, e_out = int(out_np[-1]), int(out_ep[-1])
        staged = torch.cat([ids_h, out_np, out_ep]).pin_memory().to(self.device, non_blocking=True)
        ids_d, np_d, ep_d = staged[:b], staged[b:2 * b + 1], staged[2 * b + 1:]
        x = torch.empty(n_out, self.num_features, dtype=torch.float32, device=self.device)
        ei = torch.empty(2, e_out, dtype=torch.long, device=self.device)
        batch = torch.empty(n_out, dtype=torch.long, device=self.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        gda.collate_graphs(p(self.x_all), self.num_features, p(self.ei_all), self.ei_all.size(1), p(self.node_ptr),
                           p(self.edge_ptr), p(ids_d), b, p(np_d), p(ep_d), n_out, e_out, p(x), p(ei), p(batch),
                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return Batch(x=x, edge_index=ei, y=self.y_all.index_select(0, ids_d), batch=batch, num_graphs=b, ptr=np_d)


class DataLoader:
    """``DataLoader(dataset, batch_size, shuffle)`` over a sequence of graphs.  With ``device`` set to a CUDA device
    the dataset is moved there once (``DeviceGraphDataset``) and every batch is collated on the GPU; the shuffle
    order comes from torch's own ``RandomSampler`` on the CPU generator, exactly as PyG's does."""

    def __init__(self, dataset, batch_size=1, shuffle=False, device=None, **kw):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle
        self.resident = None
        if device is not None and torch.device(device).type == "cuda" and len(dataset) > 0:
            self.resident = dataset if isinstance(dataset, DeviceGraphDataset) else DeviceGraphDataset(dataset, device)

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        # PyG's DataLoader IS torch.utils.data.DataLoader with a graph collate function: the index batches come from
        # torch's own sampler machinery, so the same torch seed gives the same batches as the reference
        # (RandomSampler draws its seed from the global CPU generator, after the iterator's base seed)
        import torch.utils.data as tud
        index_batches = tud.DataLoader(range(len(self.dataset)), batch_size=self.batch_size, shuffle=self.shuffle,
                                       collate_fn=list)
        for ids in index_batches:
            if self.resident is not None:
                yield self.resident.collate(ids)
            else:
                yield Batch.from_data_list([self.dataset[i] for i in ids])
