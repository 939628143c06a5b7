"""Multi-GPU plumbing of the junction-extraction path: one process (one ipx_handle) per GPU.

The record stream is cut into contiguous, ordered shards (global stream indexes preserved through ipx_batch.seq_base).
Every rank extracts and reduces its shard locally; partial junction rows and read-offset rows are then exchanged by
key range and folded by the owning rank.  Two transports:
  mailboxes   (GPUs of one node) the exporting kernels store the rows of junctions another rank owns straight into
              that rank's memory over NVLink (ipx_mailbox_*, ipx_partial_export_mail) -- a rank's own share never
              moves -- and publish an epoch into the slot headers; the importing side waits for the epochs on the
              device.  No send buffers, no size exchange, no collective, no host round trip.
  all-to-all  packed rows through torch.distributed (NCCL, or gloo on CPU tensors in the tests); used when no
              mailboxes are connected.  With the canon filter the ranks first agree
on each junction's creating event (an all-gather of J-sized threshold rows) and count their kept events against it.
Payloads are J-sized (MBs), so the exchange is latency-, not bandwidth-bound.

The functions come in two flavours that share everything but the transport:
  *_dist    torch.distributed (NCCL on GPUs; the byte routing also runs over gloo on CPU tensors in the tests)
  *_local   several handles in one process (tests that emulate ranks on one GPU)
This module is plumbing: no junction arithmetic happens here.
"""
import numpy as np
import torch


def key_splitters(names, n_parts):
    """Contig-aligned splitters of the key space, which is ordered by chromosome NAME (std::string::operator<), start,
    end: part p owns the names ranked [p*n/n_parts, (p+1)*n/n_parts).  Returns (split_tid, split_start): the first key of
    parts 1 .. n_parts-1."""
    order = sorted(range(len(names)), key=lambda i: names[i].encode())
    n = len(names)
    split_tid = []
    for p in range(1, n_parts):
        k = min(n - 1, (n * p) // n_parts) if n else 0
        split_tid.append(order[k] if n else 0)
    return split_tid, [0] * (n_parts - 1)


def balanced_splitters(names, samples, n_parts):
    """Splitters that give every part about the same number of junctions: `samples` = (target index, intron start)
    arrays of junction keys (every rank's ipx_sample_keys, concatenated).  Keys are ordered by chromosome NAME, start."""
    tid = np.concatenate([np.asarray(t, dtype=np.int64) for t, _ in samples]) if samples else np.zeros(0, np.int64)
    i0 = np.concatenate([np.asarray(s, dtype=np.int64) for _, s in samples]) if samples else np.zeros(0, np.int64)
    if tid.size < n_parts:
        return key_splitters(names, n_parts)
    order = sorted(range(len(names)), key=lambda i: names[i].encode())
    rank = np.zeros(len(names), np.int64)
    rank[order] = np.arange(len(names))
    key = (rank[tid] << 32) | i0
    key.sort()
    split_tid, split_start = [], []
    for p in range(1, n_parts):
        k = int(key[(key.size * p) // n_parts])
        split_tid.append(order[k >> 32])
        split_start.append(k & 0xFFFFFFFF)
    return split_tid, split_start


def balanced_splitters_dist(eng, names, device, group=None, max_keys=1 << 14):
    """Every rank samples its junction keys, the samples are all-gathered, every rank derives the same splitters."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    tid, i0 = eng.sample_keys(stride=1, max_keys=1 << 22)
    tid, i0 = tid[tid >= 0], i0[tid >= 0]  # (ids no junction took come back as -1)
    if tid.size > max_keys:
        step = (tid.size + max_keys - 1) // max_keys
        tid, i0 = tid[::step], i0[::step]
    mine = torch.full((2, max_keys), -1, dtype=torch.int64, device=device)
    mine[0, :tid.size] = torch.from_numpy(tid.astype(np.int64)).to(device)
    mine[1, :i0.size] = torch.from_numpy(i0.astype(np.int64)).to(device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    samples = []
    for g in gathered:
        a = g.cpu().numpy()
        keep = a[0] >= 0
        samples.append((a[0][keep], a[1][keep]))
    return balanced_splitters(names, samples, world)


def _engine_stream(eng, device):
    if device.type != "cuda" or not hasattr(eng, "stream"):
        return None
    if getattr(eng, "_torch_stream", None) is None:
        eng._torch_stream = torch.cuda.ExternalStream(eng.stream(), device=device)
    return eng._torch_stream


def _torch_after_engine(eng, device):
    """The engine launches on its own stream and its export calls return without waiting for the device: make torch's
    current stream (collectives, copies) wait for what the engine has enqueued.  Device-side ordering, no host sync."""
    es = _engine_stream(eng, device)
    if es is not None:
        torch.cuda.current_stream(device).wait_stream(es)


def _engine_after_torch(eng, device):
    """... and the other way round before the engine reads buffers torch's stream is still filling."""
    es = _engine_stream(eng, device)
    if es is not None:
        es.wait_stream(torch.cuda.current_stream(device))


def _byte_tensor(n_bytes, device):
    return torch.empty(max(1, n_bytes), dtype=torch.uint8, device=device)


def export_partials(eng, names, n_parts, device):
    """Finalize the local table and pack its rows destination-major.  Returns ((junction bytes, bytes per part),
    (offset-row bytes, bytes per part)); the local table is empty afterwards."""
    split_tid, split_start = key_splitters(names, n_parts)
    jrows, hrows = eng.partial_sizes(n_parts, split_tid, split_start)
    jb, hb = eng.partial_row_bytes()
    send_j = _byte_tensor(sum(jrows) * jb, device)
    send_h = _byte_tensor(sum(hrows) * hb, device)
    _engine_after_torch(eng, device)  # the allocator may hand out memory torch's stream was still using
    eng.partial_export(send_j.data_ptr(), send_h.data_ptr())
    _torch_after_engine(eng, device)
    return (send_j, [x * jb for x in jrows]), (send_h, [x * hb for x in hrows])


def import_partials(eng, recv_j, bytes_j, recv_h, bytes_h):
    """Fold what every source sent: recv_* hold the sources' segments back to back (bytes_* per source).  One import
    per source, because offset rows name their junction by its index within the same source's junction rows."""
    jb, hb = eng.partial_row_bytes()
    oj = oh = 0
    for nj, nh in zip(bytes_j, bytes_h):
        if nj or nh:
            eng.partial_import(recv_j.data_ptr() + oj, nj // jb, recv_h.data_ptr() + oh, nh // hb)
        oj += nj
        oh += nh


def exchange_counts(count_lists, device, group=None):
    """What every source will send here: one small all-to-all for all kinds of rows.  count_lists: per kind, the bytes
    this rank sends to each destination; returns per kind the bytes it receives from each source."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = torch.tensor(count_lists, dtype=torch.int64, device=device).t().contiguous()  # [dest][kind]
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts.view(-1), counts.view(-1), group=group)
    rc = recv_counts.cpu().tolist()  # [source][kind]
    return [[int(rc[s][k]) for s in range(world)] for k in range(len(count_lists))]


def all_to_all_bytes(send, send_counts, recv_counts, group=None):
    """Variable-size all-to-all of a uint8 tensor laid out destination-major; returns the received bytes in
    source-rank order."""
    import torch.distributed as dist
    recv = torch.empty(max(1, sum(recv_counts)), dtype=torch.uint8, device=send.device)
    dist.all_to_all_single(recv[:sum(recv_counts)], send[:sum(send_counts)], list(recv_counts), list(send_counts), group=group)
    return recv[:sum(recv_counts)]


def connect_mailboxes_dist(eng, device, junction_rows_per_source, hist_rows_per_source, group=None):
    """Create this rank's mailbox and map those of all other ranks of the node (CUDA IPC handles, all-gathered)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    eng.mailbox_create(world, rank, junction_rows_per_source, hist_rows_per_source)
    mine = torch.frombuffer(bytearray(eng.mailbox_ipc_handle()), dtype=torch.uint8).to(device)
    handles = [torch.empty(64, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(handles, mine, group=group)
    eng.mailbox_connect_ipc([bytes(h.cpu().numpy().tobytes()) for h in handles])
    dist.barrier(group=group)


def connect_mailboxes_local(engines, junction_rows_per_source, hist_rows_per_source):
    """The same for handles of one process (emulated ranks on one GPU): plain device pointers."""
    world = len(engines)
    for r, e in enumerate(engines):
        e.mailbox_create(world, r, junction_rows_per_source, hist_rows_per_source)
    ptrs = [e.mailbox_ptr() for e in engines]
    for e in engines:
        e.mailbox_connect_ptrs(ptrs)


def exchange_mail_dist(eng, names, device, group=None, splitters=None):
    """One key-range exchange through the mailboxes: export what other ranks own into their memory, then fold what
    arrived here -- the importing kernels wait on the device for every source's epoch.  Everything is enqueued on the
    engine's stream; there is no collective and nothing waits on the host."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    split_tid, split_start = splitters if splitters is not None else key_splitters(names, world)
    eng.partial_export_mail(split_tid, split_start)
    _debug_sync("export")
    eng.partial_import_mail()
    _debug_sync("import")


def _debug_sync(what):
    """IPX_DEBUG_SYNC=export,import: wait for the device after that phase (fault hunting; never set in measurements)."""
    import os
    if what in os.environ.get("IPX_DEBUG_SYNC", "").split(","):
        torch.cuda.synchronize()


def settle_exchange_dist(eng, names, device, group=None, splitters=None):
    """Before ipx_flush: did any rank fold events after its export (its junction table had been full and was grown)?
    Then everybody repeats the exchange -- only what is still held for other ranks moves."""
    import torch.distributed as dist
    for _ in range(8):
        flag = torch.tensor([1 if eng.exchange_settle() else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) == 0:
            return
        exchange_mail_dist(eng, names, device, group, splitters)
    raise RuntimeError("the exchange did not settle")


def exchange_mail_local(engines, names, device, splitters=None):
    split_tid, split_start = splitters if splitters is not None else key_splitters(names, len(engines))
    for _ in range(8):
        for e in engines:
            e.partial_export_mail(split_tid, split_start)
        # the handles have separate streams and the importers wait on the device for every exporter's epoch: all the
        # exports must be enqueued before the first import, or an importer could wait for a kernel queued behind it
        for e in engines:
            e.partial_import_mail()
        if not any([e.exchange_settle() for e in engines]):
            return
    raise RuntimeError("the exchange did not settle")


def exchange_dist(eng, names, device, group=None, splitters=None):
    """One key-range exchange over torch.distributed; returns the bytes this rank sent + received (None through
    mailboxes, where nobody counts them on the host)."""
    import torch.distributed as dist
    if getattr(eng, "mailbox_connected", False):
        return exchange_mail_dist(eng, names, device, group, splitters)
    world = dist.get_world_size(group)
    (send_j, cj), (send_h, ch) = export_partials(eng, names, world, device)
    rj, rh = exchange_counts([cj, ch], device, group)
    recv_j = all_to_all_bytes(send_j, cj, rj, group)
    recv_h = all_to_all_bytes(send_h, ch, rh, group)
    _engine_after_torch(eng, device)  # the collectives run on torch's stream: the rows must have landed
    import_partials(eng, recv_j, rj, recv_h, rh)
    _torch_after_engine(eng, device)  # the receive buffers go back to torch's allocator: not before the import has read them
    return sum(cj) + sum(ch) + sum(rj) + sum(rh)


def exchange_local(engines, names, device):
    """The same exchange among handles of one process (rank = position in `engines`)."""
    if all(getattr(e, "mailbox_connected", False) for e in engines):
        return exchange_mail_local(engines, names, device)
    world = len(engines)
    exports = [export_partials(e, names, world, device) for e in engines]
    moved = 0
    for dst, eng in enumerate(engines):
        pieces_j, pieces_h = [], []
        for src in range(world):
            (sj, cj), (sh, ch) = exports[src]
            oj, oh = sum(cj[:dst]), sum(ch[:dst])
            pieces_j.append(sj[oj:oj + cj[dst]])
            pieces_h.append(sh[oh:oh + ch[dst]])
        rj, rh = torch.cat(pieces_j), torch.cat(pieces_h)
        moved += rj.numel() + rh.numel()
        _engine_after_torch(eng, device)
        import_partials(eng, rj, [p.numel() for p in pieces_j], rh, [p.numel() for p in pieces_h])
        _torch_after_engine(eng, device)
    return moved


def _export_thresholds(eng, device):
    n, row_bytes = eng.thresholds_rows()
    rows = _byte_tensor(n * row_bytes, device)
    eng.thresholds_export(rows.data_ptr())
    return rows[:n * row_bytes], row_bytes


def agree_thresholds_dist(eng, device, group=None):
    """Canon filter: every rank learns the earliest creating event of each junction it has seen, then counts the
    junction events it kept (ipx_params.deferred_counting) against it."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows, row_bytes = _export_thresholds(eng, device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([rows.numel()], dtype=torch.int64, device=device), group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(1, max(sizes))
    padded = torch.zeros(cap, dtype=torch.uint8, device=device)
    padded[:rows.numel()] = rows
    gathered = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    _engine_after_torch(eng, device)
    for r in range(world):
        if r != dist.get_rank(group) and sizes[r]:
            eng.thresholds_import(gathered[r].data_ptr(), sizes[r] // row_bytes)
    eng.count_deferred()
    return sum(sizes)


def agree_thresholds_local(engines, device):
    exported = [_export_thresholds(e, device) for e in engines]
    for i, eng in enumerate(engines):
        for j, (rows, row_bytes) in enumerate(exported):
            if j != i and rows.numel():
                eng.thresholds_import(rows.data_ptr(), rows.numel() // row_bytes)
    for eng in engines:
        eng.count_deferred()


def global_numbering(creation_seqs):
    """sjN over several tables: the rank of every junction's creation_seq among all of them (creation_seq values are
    global stream positions, so they are distinct).  Takes and returns one array per table."""
    if not creation_seqs:
        return []
    allv = np.sort(np.concatenate([np.asarray(c, dtype=np.uint64) for c in creation_seqs]))
    return [np.searchsorted(allv, np.asarray(c, dtype=np.uint64)).astype(np.uint32) for c in creation_seqs]


class MergeRowsPart:
    """The rows of an ipx_merge_rows that fall into one key range, as a new ipx_merge_rows over numpy copies."""

    def __init__(self, rows, keep):
        import ctypes as C
        from ._native import MergeRows
        n = int(rows.n_rows)

        def col(ptr, ctype, width=1):
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n * width,)) if n else np.zeros(0, ctype)
            return a.reshape(n, width) if width > 1 else a

        self.cols = {
            "chrom": np.ascontiguousarray(col(rows.chrom, C.c_int32)[keep]),
            "intron_start": np.ascontiguousarray(col(rows.intron_start, C.c_uint32)[keep]),
            "intron_end": np.ascontiguousarray(col(rows.intron_end, C.c_uint32)[keep]),
            "anchor_start": np.ascontiguousarray(col(rows.anchor_start, C.c_uint32)[keep]),
            "anchor_end": np.ascontiguousarray(col(rows.anchor_end, C.c_uint32)[keep]),
            "strand": np.ascontiguousarray(col(rows.strand, C.c_uint8)[keep]),
            "sites": np.ascontiguousarray(col(rows.sites, C.c_uint32)[keep]),
            "counts": np.ascontiguousarray(col(rows.counts, C.c_uint32, 3)[keep]),
            "confidence": np.ascontiguousarray(col(rows.confidence, C.c_float)[keep]),
        }
        m = MergeRows()
        m.n_rows = int(self.cols["chrom"].shape[0])
        for k, a in self.cols.items():
            setattr(m, k, a.ctypes.data)
        m.site_strings = rows.site_strings
        m.n_site_strings = rows.n_site_strings
        self.struct = m


def merge_rows_for_part(rows, names, n_parts, part):
    """intronProspectorMerge across GPUs: keys are sharded by chromosome range BEFORE the fold, so every row of a
    junction reaches the same rank in (file, row) order and no exchange is needed.  Returns the rows of `part`."""
    import ctypes as C
    n = int(rows.n_rows)
    chrom = np.ctypeslib.as_array(C.cast(rows.chrom, C.POINTER(C.c_int32)), shape=(n,)) if n else np.zeros(0, np.int32)
    order = sorted(range(len(names)), key=lambda i: names[i].encode())
    rank = np.zeros(len(names), np.int64)
    rank[order] = np.arange(len(names))
    lo = (len(names) * part) // n_parts
    hi = (len(names) * (part + 1)) // n_parts
    r = rank[chrom]
    return MergeRowsPart(rows, (r >= lo) & (r < hi))
