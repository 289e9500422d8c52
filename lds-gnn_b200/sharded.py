"""Row-block sharded direct outer step: theta, A_tilde and every row-local buffer are partitioned
by contiguous row blocks over the ranks of one box. theta / A_tilde never move:

  * sampling needs no exchange — the Philox draw of edge (i, j) is keyed on (min, max), so the owners of row i and of
    row j regenerate the same bit, and the K3 update is symmetric bit for bit (csrc/lds_k3_theta_update.cu);
  * each of the four propagations Z_rows = r_rows * (A_tilde_rows @ (r * P)) needs the scaled N x w operand of ALL
    rows: one all-gather per propagation — the only data-path collective. Every rank leaves its rows already in the
    layout the tensor cores read (K-major bf16 hi/lo block [2][hp][rows per rank]); the gathered array
    [rank][2][hp][rows per rank] is read through a 3-D TMA tensor map, so nothing is re-laid out after the gather
    (16.8 MB at N = 65 536 for the two h = 64 wide operands, 4.2 MB for the two class-wide ones);
  * the closed-form theta update needs the factor rows of ALL nodes and c [N]: for the tensor-core SGD update these
    are the packed bf16 rows the BWD1 epilogue writes (8 B per factor element, lds_k3.cuh), gathered as they are;
    for the CUDA-core update (Adam) the fp32 rows fa, fb [N, d];
  * loss / accuracy: one all-reduce of two floats.

`ShardedOuterStep` drives one rank through the phases of `lds_outer_step` (include/lds_b200.h, LDS_PHASE_*);
the exchange object is either `DistComm` (torch.distributed: NCCL over NVLink on GPUs) or, for single-GPU tests,
`run_local_group`, which steps several shards of one matrix through the same phases and gathers by concatenation.
"""
import torch

from . import _lib
from . import kernels as K

PHASES = (_lib.PHASE_LAYER1, _lib.PHASE_LAYER2, _lib.PHASE_BWD2, _lib.PHASE_BWD1)


def shard_bounds(n, world, rank):
    """Contiguous row block of `rank`; block starts are multiples of 128 (K2 row panels, even for K1's row pairs)."""
    per = -(-n // world)
    per = -(-per // 128) * 128
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi - lo


class DistComm:
    """Exchange over torch.distributed (backend nccl on GPUs, gloo in CPU tests). Shards may be ragged."""

    def __init__(self, n, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.bounds = [shard_bounds(n, self.world, r) for r in range(self.world)]
        self.equal = len({b[1] for b in self.bounds}) == 1

    def all_gather_rows(self, local, out):
        """local [rows_r, w] -> out [n, w] (row blocks in rank order)."""
        if self.equal:
            self.dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
            return out
        per = max(cnt for _, cnt in self.bounds)           # ragged last block: pad to the common block size
        padded = local.new_zeros((per,) + tuple(local.shape[1:]))
        padded[:local.shape[0]] = local
        gathered = local.new_empty((self.world * per,) + tuple(local.shape[1:]))
        self.dist.all_gather_into_tensor(gathered, padded, group=self.group)
        for r, (lo, cnt) in enumerate(self.bounds):
            out[lo:lo + cnt] = gathered[r * per:r * per + cnt]
        return out

    def all_gather_flat(self, local, out):
        """Equal-size contiguous chunks: out = concat over ranks of `local`."""
        self.dist.all_gather_into_tensor(out, local, group=self.group)
        return out

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, group=self.group)
        return t

    # ---- gather buffers: `gather_into` fills gb.t[offset + r * chunk : offset + (r + 1) * chunk] with rank r's `local`
    def gather_buffer(self, numel, dtype, device):
        return GatherBuffer(torch.zeros(numel, dtype=dtype, device=device))

    def gather_into(self, local, gb, offset, chunk, count=None, barrier=True):
        """`local` holds `count` (default `chunk`) valid elements; equal `chunk` strides between ranks. `barrier` only matters for
        the peer-memory exchange (several pushes may share the barrier of the last one)."""
        if count is not None and count != chunk:
            padded = local.new_zeros(chunk)
            padded[:count] = local[:count]
            local = padded
        self.dist.all_gather_into_tensor(gb.t[offset:offset + self.world * chunk], local[:chunk], group=self.group)


class GatherBuffer:
    def __init__(self, tensor, ctx=None):
        self.t, self.ctx = tensor, ctx


class LocalComm:
    """The exchange of a single rank that owns every row (world size 1, no process group): gathers are copies."""

    def __init__(self, n):
        self.world, self.rank = 1, 0
        self.bounds = [shard_bounds(n, 1, 0)]
        self.equal = True

    def all_gather_rows(self, local, out):
        out[:local.shape[0]].copy_(local)
        return out

    def all_reduce_sum(self, t):
        return t

    def gather_buffer(self, numel, dtype, device):
        return GatherBuffer(torch.zeros(numel, dtype=dtype, device=device))

    def gather_into(self, local, gb, offset, chunk, count=None, barrier=True):
        k = chunk if count is None else count
        gb.t[offset:offset + k].copy_(local.reshape(-1)[:k])


def make_comm(n, device):
    """The exchange object of this process: NVLink peer-memory push over symmetric memory when torch.distributed runs NCCL on
    CUDA devices (falls back to NCCL all-gathers), gloo / other backends through torch.distributed, a plain copy without a
    process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return LocalComm(n)
    if device.type == "cuda" and dist.get_backend() == "nccl":
        try:
            return SymmComm(n)
        except Exception:                                       # noqa: BLE001  (symmetric memory unavailable: NCCL all-gathers)
            pass
    return DistComm(n)


class SymmComm(DistComm):
    """Exchange over NVLink peer memory instead of NCCL: the gather buffers are torch symmetric memory (every rank maps
    every peer's buffer), `gather_into` is ONE kernel of ours that stores this rank's block into all peers' buffers with
    128-bit stores (csrc/lds_common.cu: peer_push_kernel) followed by the signal-pad barrier of the symmetric-memory
    handle. At the sizes of this path (0.5 - 5 MB per rank) an NCCL all-gather is dominated by its launch and protocol
    latency (~35-130 us measured, 7 per step); the push runs at NVLink store bandwidth."""

    def __init__(self, n, group=None):
        super().__init__(n, group)
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        pg = group if group is not None else self.dist.group.WORLD
        self.group_name = pg.group_name
        self.lib = _lib.load()

    def gather_buffer(self, numel, dtype, device):
        t = self.symm_mem.empty(int(numel), dtype=dtype, device=device)
        t.zero_()
        hdl = self.symm_mem.rendezvous(t, self.group_name)
        ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize()
        hdl.barrier(channel=0)                              # everybody's buffer is zeroed before anybody pushes
        return GatherBuffer(t, (hdl, ptrs))

    def gather_into(self, local, gb, offset, chunk, count=None, barrier=True):
        hdl, ptrs = gb.ctx
        esz = local.element_size()
        nbytes = -(-(chunk if count is None else count) * esz // 16) * 16
        _lib.check(self.lib.lds_peer_push(local.data_ptr(), ptrs.data_ptr(), self.world, (offset + self.rank * chunk) * esz, nbytes,
                                          K._stream()), "lds_peer_push")
        if barrier:                                          # pushes issued before this one on the same stream are covered too
            hdl.barrier(channel=0)


class ShardedOuterStep:
    def __init__(self, n, row0, rows, x_local, y_local, mask_local, mask_count, hidden, classes, sparse_features=None):
        self.n, self.row0, self.rows = int(n), int(row0), int(rows)
        self.h, self.c = int(hidden), int(classes)
        self.eng = K.OuterStep(n, x_local, y_local, mask_local, hidden, classes, sparse_features=sparse_features,
                               row0=row0, rows=rows, mask_count=mask_count)
        dev = x_local.device
        w = max(self.h, self.c)
        self.ldf = int(_lib.load().lds_outer_step_factor_ld(self.h, self.c))
        self.opnd_full = torch.empty((self.n, w), dtype=torch.float32, device=dev)          # legacy fp32 row exchange (packed=False)
        self.packed = True
        self.world_rows = None                                 # rows per rank (multiple of 128), set by the driver / comm
        self.fa_full = self.fb_full = None                   # fp32 factor rows: only for the CUDA-core update (allocated on first use)
        self.kf = int(_lib.load().lds_outer_step_packed_k(self.h, self.c))
        self.f_full = torch.empty((self.n, self.kf), dtype=K.BF16, device=dev)
        self.c_full = torch.empty(self.n, dtype=torch.float32, device=dev)

    def set_weights(self, *w):
        self.eng.set_weights(*w)

    def _setup_packed(self, per):
        """Buffers of the packed operand exchange: send block [2][hp][per] bf16 and the gathered [R][2][hp][per]."""
        if self.world_rows == per:
            return
        self.world_rows = int(per)
        lib = _lib.load()
        self.hp = {ph: int(lib.lds_outer_step_operand_hp(self.h, self.c, ph)) for ph in PHASES}
        hpmax = max(self.hp.values())
        nblk = -(-self.n // per)
        dev = self.opnd_full.device
        raw = torch.zeros(2 * hpmax * per + 512, dtype=K.BF16, device=dev)       # rows beyond this rank's count stay zero
        off = ((-raw.data_ptr()) % 1024) // 2
        self.send = raw[off:off + 2 * hpmax * per]
        self._send_raw = raw
        self.recv = torch.zeros(nblk * 2 * hpmax * per, dtype=K.BF16, device=dev)

    def phase(self, theta_local, phases, **kw):
        if self.packed and self.world_rows:
            return self.eng.run(theta_local, phases=phases, opnd_full=self.recv, opnd_send=self.send, opnd_rank_rows=self.world_rows,
                                fa_full=self.fa_full, fb_full=self.fb_full, c_full=self.c_full, f_full=self.f_full, **kw)
        return self.eng.run(theta_local, phases=phases, opnd_full=self.opnd_full, fa_full=self.fa_full, fb_full=self.fb_full,
                            c_full=self.c_full, f_full=self.f_full, **kw)

    @staticmethod
    def tensor_core_update(kw):
        return kw.get("opt_kind", _lib.OPT_SGD) == _lib.OPT_SGD and not (kw.get("k3_flags", 0) & _lib.K3_SIMT)

    def factor_buffers(self, kw):
        """(workspace buffer name, gathered tensor) pairs PHASE_UPDATE needs besides c."""
        if self.tensor_core_update(kw):
            return [("fpack", self.f_full)]
        if self.fa_full is None:
            dev = self.opnd_full.device
            self.fa_full = torch.empty((self.n, self.ldf), dtype=torch.float32, device=dev)
            self.fb_full = torch.empty((self.n, self.ldf), dtype=torch.float32, device=dev)
        return [("fa", self.fa_full), ("fb", self.fb_full)]

    def _setup_comm(self, comm):
        """Gather buffers owned by the exchange object (plain tensors for NCCL / gloo, symmetric memory for SymmComm)."""
        if getattr(self, "_comm", None) is comm:
            return
        per = comm.bounds[0][1]
        self._setup_packed(per)
        dev = self.opnd_full.device
        hpmax = max(self.hp.values())
        self._half = comm.world * 2 * hpmax * per
        self.g_opnd = comm.gather_buffer(2 * self._half, K.BF16, dev)        # double-buffered: exchange e fills half e & 1
        self.g_f = comm.gather_buffer(comm.world * per * self.kf, K.BF16, dev)
        self.g_c = comm.gather_buffer(comm.world * per + 4, torch.float32, dev)
        self.g_s = comm.gather_buffer(comm.world * 4, torch.float32, dev)     # per-rank (loss, acc) partial sums: they ride with the update's exchange
        self.f_full = self.g_f.t.view(-1, self.kf)
        self.c_full = self.g_c.t
        self._xchg = 0
        self._comm = comm

    def run(self, theta_local, comm, lr, seed, step, dropout_p=0.0, update=True, **kw):
        """One sharded outer step on this rank. Returns a device tensor (loss, acc) of the WHOLE graph."""
        kw = dict(lr=lr, seed=seed, step=step, dropout_p=dropout_p, update=update, **kw)
        if not self.packed:
            return self._run_legacy(theta_local, comm, kw)
        self._setup_comm(comm)
        per = self.world_rows
        self.phase(theta_local, _lib.PHASE_SAMPLE, **kw)
        for ph in PHASES:
            off = (self._xchg & 1) * self._half
            comm.gather_into(self.send, self.g_opnd, off, 2 * self.hp[ph] * per)
            self.recv = self.g_opnd.t[off:off + self._half]
            self._xchg += 1
            self.phase(theta_local, ph, **kw)
        # factor rows, the ranks' (loss, acc) partial sums and c leave back to back and share ONE barrier (three pushes, no
        # all-reduce: a 2-float NCCL all-reduce costs more than the whole exchange of c)
        if self.tensor_core_update(kw):
            comm.gather_into(self.eng.buffer("fpack").reshape(-1), self.g_f, 0, per * self.kf, count=self.rows * self.kf, barrier=False)
        else:
            for name, full in self.factor_buffers(kw):
                comm.all_gather_rows(self.eng.buffer(name), full)
        comm.gather_into(self.eng.scalars, self.g_s, 0, 4, barrier=False)
        comm.gather_into(self.eng.buffer("cvec"), self.g_c, 0, per, count=self.rows)
        self.phase(theta_local, _lib.PHASE_UPDATE, **kw)
        return self.g_s.t.view(-1, 4)[:, :2].sum(dim=0)

    def _run_legacy(self, theta_local, comm, kw):
        """fp32 row exchange + re-layout kernel per propagation (kept for comparison / tests)."""
        self.phase(theta_local, _lib.PHASE_SAMPLE, **kw)
        for ph in PHASES:
            comm.all_gather_rows(self.eng.buffer("operand"), self.opnd_full)
            self.phase(theta_local, ph, **kw)
        scalars = comm.all_reduce_sum(self.eng.scalars[:2].clone())
        for name, full in self.factor_buffers(kw):
            comm.all_gather_rows(self.eng.buffer(name), full)
        comm.all_gather_rows(self.eng.buffer("cvec").view(-1, 1), self.c_full.view(-1, 1))
        self.phase(theta_local, _lib.PHASE_UPDATE, **kw)
        return scalars


def run_local_group(shards, thetas, lr, seed, step, dropout_p=0.0, update=True, **kw):
    """Single-GPU emulation of the multi-rank step (tests): every shard goes through the same phases, the exchange is a
    concatenation. `shards`: ShardedOuterStep objects covering [0, n) in order; `thetas`: their local theta row blocks."""
    kw = dict(lr=lr, seed=seed, step=step, dropout_p=dropout_p, update=update, **kw)

    def gather(name, attr, view=None):
        parts = [s.eng.buffer(name) for s in shards]
        full = torch.cat([p if view is None else p.view(*view) for p in parts], dim=0)
        for s in shards:
            getattr(s, attr).view(full.shape).copy_(full)

    for s in shards:
        s.factor_buffers(kw)                                 # allocate the fp32 factor buffers if this step needs them

    per = max(s.rows for s in shards)
    per = -(-per // 128) * 128
    for s in shards:
        if s.packed:
            s._setup_packed(per)
    for s, t in zip(shards, thetas):
        s.phase(t, _lib.PHASE_SAMPLE, **kw)
    for ph in PHASES:
        if shards[0].packed:
            cnt = 2 * shards[0].hp[ph] * per
            full = torch.cat([s.send[:cnt] for s in shards])
            for s in shards:
                s.recv[:full.numel()].copy_(full)
        else:
            gather("operand", "opnd_full")
        for s, t in zip(shards, thetas):
            s.phase(t, ph, **kw)
    scalars = sum(s.eng.scalars[:2].clone() for s in shards)
    if shards[0].tensor_core_update(kw):
        gather("fpack", "f_full")
    else:
        gather("fa", "fa_full")
        gather("fb", "fb_full")
    gather("cvec", "c_full", view=(-1,))
    for s, t in zip(shards, thetas):
        s.phase(t, _lib.PHASE_UPDATE, **kw)
    return scalars
