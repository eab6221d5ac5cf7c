"""One rank of the time-sharded flagship chain: streams, rings and the per-block
schedule.  One process per GPU (torch.distributed, NCCL); nothing here computes, the data path is libgr_cuda.

Per block b = step * world + rank a rank runs

  main stream   front(b): channelizer -> discriminator + matched filter on its block + halo (finite memory: no
                dependence on any other rank), strictly after the rank's own previous clock-recovery kernel
  halo stream   NCCL isend/irecv of the input halo of block b + world, posted one step ahead
  tail stream   recv(loop state of block b-1) -> clock-recovery kernel -> isend(loop state of block b).  These kernels
                form the ONE serial chain over all blocks of all ranks; the kernel reads the state from the receive
                buffer and writes it to the send buffer itself (no copies on the chain)
  corr stream   recv(correlator registers of b-1) -> time-parallel correlator of block b -> isend: a second, much
                shorter chain that trails the first one

and the sync hits of all its blocks accumulate on the device (absolute bit indices) until `gather_hits` brings every
rank's list to rank 0 with NCCL send/recv (the "final result gather").
"""
import os
import sys
import time

import torch
import torch.distributed as dist


def _debug():
    return bool(os.environ.get("GRB_SHARD_DEBUG"))


def _dbg(msg):
    if os.environ.get("GRB_SHARD_DEBUG"):
        sys.stderr.write("[shard rank %s %.2f] %s\n" % (os.environ.get("RANK", "0"), time.time() % 1000, msg))
        sys.stderr.flush()


class _Ptr:
    """A raw device address with the .data_ptr() the chain wrapper asks tensors for."""

    def __init__(self, addr):
        self.addr = int(addr)

    def data_ptr(self):
        return self.addr


class PeerRing:
    """The loop-state hand-off between neighbouring ranks WITHOUT NCCL kernels: every rank owns
    one device allocation [flags | clock-recovery state | correlator registers], exports it with CUDA IPC, and opens its
    right neighbour's.  The tail kernel writes its final state STRAIGHT INTO the right neighbour's buffer (peer stores over
    NVLink), a stream memory operation (cuStreamWriteValue32, which fences system-wide first) then publishes the block
    number in the neighbour's flag, and the neighbour's tail stream holds a cuStreamWaitValue32(flag >= block) in front of
    its kernel.  No receive has to be posted, nothing spins on an SM, nothing needs a matching call: the ordering problems
    of the send/recv version (prepost, eager communicator creation for these two rings) do not exist here."""

    FLAG_BYTES = 256

    def __init__(self, plan, device, mm_bytes, corr_bytes):
        from cuda.bindings import driver as cu
        self.cu, self.plan = cu, plan
        torch.cuda.synchronize(device)                       # torch's primary context is current on this thread
        self.mm_off = self.FLAG_BYTES
        self.corr_off = self.mm_off + ((mm_bytes + 255) // 256) * 256
        self.nbytes = self.corr_off + ((corr_bytes + 255) // 256) * 256
        self._ck(cu.cuInit(0))
        self.base = int(self._ck(cu.cuMemAlloc(self.nbytes)))
        self._ck(cu.cuMemsetD8(self.base, 0, self.nbytes))
        handle = self._ck(cu.cuIpcGetMemHandle(self.base))
        mine = torch.tensor(list(bytes(handle.reserved)), dtype=torch.uint8, device=device)
        every = [torch.zeros_like(mine) for _ in range(plan.world)]
        dist.all_gather(every, mine)
        h = cu.CUipcMemHandle()
        h.reserved = bytes(every[plan.right].cpu().tolist())
        self.peer = int(self._ck(cu.cuIpcOpenMemHandle(h, cu.CUipcMem_flags.CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS)))
        torch.cuda.synchronize(device)
        # (no collective from here on: the caller's all_reduce is the next one on EVERY rank, whether this constructor
        # succeeded or raised after the all_gather)

    def _ck(self, res):
        err = res[0]
        if int(err) != 0:
            raise RuntimeError("CUDA driver call failed: %r" % (err,))
        return res[1] if len(res) == 2 else (res[1:] if len(res) > 2 else None)

    # own receive buffers / the right neighbour's, as pointers the chain takes
    def mm_in(self):
        return _Ptr(self.base + self.mm_off)

    def corr_in(self):
        return _Ptr(self.base + self.corr_off)

    def mm_out(self):
        return _Ptr(self.peer + self.mm_off)

    def corr_out(self):
        return _Ptr(self.peer + self.corr_off)

    def wait(self, stream, which, value):
        """stream waits until own flag[which] >= value"""
        cu = self.cu
        self._ck(cu.cuStreamWaitValue32(cu.CUstream(stream.cuda_stream), self.base + 4 * which, int(value) & 0x7fffffff,
                                        cu.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ))

    def publish(self, stream, which, value):
        """(after the kernel on this stream) the right neighbour's flag[which] = value"""
        cu = self.cu
        self._ck(cu.cuStreamWriteValue32(cu.CUstream(stream.cuda_stream), self.peer + 4 * which, int(value) & 0x7fffffff, 0))

    def close(self):
        try:
            self.cu.cuIpcCloseMemHandle(self.peer)
            self.cu.cuMemFree(self.base)
        except Exception:
            pass


class ShardedChain:
    def __init__(self, ch, plan, device, rows, halo, history_rows, handoff=None):
        self.ch, self.plan, self.dev = ch, plan, device
        self.R, self.halo, self.Th = int(rows), int(halo), int(history_rows)
        w = plan.world
        assert w >= 2
        # one process group per message kind: their orders cannot interleave (created in the same order everywhere)
        self.halo_group = dist.new_group()
        self.mm_group = dist.new_group()
        self.corr_group = dist.new_group()
        self.tail_ts = torch.cuda.Stream(device=device)
        self.corr_ts = torch.cuda.Stream(device=device)
        self.halo_ts = torch.cuda.Stream(device=device)
        u8 = dict(dtype=torch.uint8, device=device)
        self.mm_in = torch.zeros(ch.mm_state_bytes(), **u8)
        self.mm_out = [torch.zeros(ch.mm_state_bytes(), **u8) for _ in range(2)]
        self.corr_in = torch.zeros(ch.corr_state_bytes(), **u8)
        self.corr_out = [torch.zeros(ch.corr_state_bytes(), **u8) for _ in range(2)]
        self.mm_send = [None, None]
        self.corr_send = [None, None]
        self.halo_ev = [None, None]
        self.front_ev = [None, None]
        self._pre = None          # (step, mm work, corr work): receives posted ahead of a device-wide synchronisation
        self._marks = []          # GRB_SHARD_DEBUG: (label, event) in launch order
        # loop-state hand-off: peer memory + stream memory operations when CUDA IPC works on this box, else NCCL send/recv
        handoff = handoff or os.environ.get("GRB_SHARD_HANDOFF", "peer")
        self.ring = None
        if handoff == "peer":
            ok = torch.zeros(1, dtype=torch.int32, device=device)
            try:
                self.ring = PeerRing(plan, device, ch.mm_state_bytes(), ch.corr_state_bytes())
                ok += 1
            except Exception as e:      # no IPC in this container, no peer access, ...: every rank must agree on the fallback
                sys.stderr.write("[shard rank %d] peer hand-off unavailable (%r): NCCL send/recv\n" % (plan.rank, e))
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if self.ring is not None:
                    self.ring.close()
                self.ring = None
        self.handoff = "peer memory + stream memory ops" if self.ring is not None else "nccl send/recv"
        ch.set_accumulate_hits(True)
        ch.clear_hits(torch.cuda.current_stream(device).cuda_stream)
        # Create every point-to-point communicator NOW, with one symmetric exchange per group: created lazily by the
        # first send / receive inside a step, the creation is a host-side rendezvous of the two ranks, and the first
        # use of each group happens at different places of the schedule on different ranks (rank 0 sends before it
        # ever receives) -- the ranks then wait for each other in different groups for ever.
        for grp in (self.halo_group, self.mm_group, self.corr_group):
            a = torch.zeros(4, dtype=torch.uint8, device=device)
            b = torch.zeros(4, dtype=torch.uint8, device=device)
            for wk in dist.batch_isend_irecv([dist.P2POp(dist.isend, a, plan.right, group=grp),
                                              dist.P2POp(dist.irecv, b, plan.left, group=grp)]):
                wk.wait()
        torch.cuda.synchronize(device)

    # ---- halo of step s into xbuf[s % 2] (tap history + warm-up rows of the left block) -----------------------------
    def post_halo(self, s, xbuf):
        p, buf = self.plan, xbuf[s % 2]
        H = self.Th + self.halo
        with torch.cuda.stream(self.halo_ts):
            if self.front_ev[s % 2] is not None:
                self.halo_ts.wait_event(self.front_ev[s % 2])      # the front of step s-2 has read this copy's head
            ops = [dist.P2POp(dist.isend, buf[buf.shape[0] - H:], p.right, group=self.halo_group),
                   dist.P2POp(dist.irecv, buf[:H], p.left, group=self.halo_group)]
            for wk in dist.batch_isend_irecv(ops):
                wk.wait()
            ev = torch.cuda.Event()
            ev.record(self.halo_ts)
            self.halo_ev[s % 2] = ev

    def step(self, s, xbuf, last_step, exchange_halo=True):
        """exchange_halo = False: the caller has put the block INCLUDING its halo rows into xbuf[s % 2] (host-fed shards:
        the halo arrives with the rank's own host-to-device copy, no NCCL exchange)."""
        ch, p = self.ch, self.plan
        cur = torch.cuda.current_stream(self.dev)
        stream = cur.cuda_stream
        _dbg("step %d: front" % s)
        if exchange_halo:
            cur.wait_event(self.halo_ev[s % 2])                     # this step's halo has landed
        ch.seek_async(p.abs_start(s) - self.halo, stream)
        # the front runs after the rank's own previous clock-recovery kernel, not underneath it: both then run at
        # their stand-alone speed, and the device has the neighbours' slots of the serial chain to wait through anyway
        cur.wait_stream(self.tail_ts)
        ch.process_front_device(xbuf[s % 2], self.halo + self.R, stream)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.front_ev[s % 2] = ev
        self._mark("front %d" % s, cur)
        _dbg("step %d: halo of the next step" % s)
        if exchange_halo:
            self.post_halo(s + 1, xbuf)                             # one exchange per step, one step ahead
        _dbg("step %d: tail" % s)
        k = s % 2
        final = s == last_step and p.rank == p.world - 1            # nobody waits for the last block's state
        if self.ring is not None:
            # block number of this step's block; the left neighbour publishes b (= its block b - 1, plus one) when the state
            # of block b - 1 is in this rank's buffer, this rank publishes b + 1 to the right
            b = p.block_index(s)
            has_left = p.has_left_state(s)
            with torch.cuda.stream(self.tail_ts):
                self.tail_ts.wait_event(ev)
                if has_left:
                    self.ring.wait(self.tail_ts, 0, b)
                self._mark("mm state in %d" % s, self.tail_ts)
                ch.process_tail_mm_device(self.ring.mm_in() if has_left else None, None if final else self.ring.mm_out(),
                                          self.tail_ts.cuda_stream)
                self._mark("mm kernel %d" % s, self.tail_ts)
                if not final:
                    self.ring.publish(self.tail_ts, 0, b + 1)
            with torch.cuda.stream(self.corr_ts):
                if has_left:
                    self.ring.wait(self.corr_ts, 1, b)
                ch.process_tail_corr_device(self.ring.corr_in() if has_left else None, None if final else self.ring.corr_out(),
                                            self.corr_ts.cuda_stream)
                self._mark("corr kernel %d" % s, self.corr_ts)
                if not final:
                    self.ring.publish(self.corr_ts, 1, b + 1)
            _dbg("step %d: done" % s)
            return
        with torch.cuda.stream(self.tail_ts):
            ts = self.tail_ts.cuda_stream
            # an NCCL receive spins on an SM until its peer sends: posted before the front is done it takes that SM
            # from the front's persistent one-CTA-per-SM FFT, which then needs a second wave
            self.tail_ts.wait_event(ev)
            if self.mm_send[k] is not None:
                self.mm_send[k].wait()                              # the send of step s-2 has left this buffer
            has_left = p.has_left_state(s)
            pre = self._pre if (self._pre is not None and self._pre[0] == s) else None
            if pre is not None:
                pre[1].wait()
            elif has_left:
                dist.recv(self.mm_in, src=p.left, group=self.mm_group)
            self._mark("mm state in %d" % s, self.tail_ts)
            ch.process_tail_mm_device(self.mm_in if has_left else None, self.mm_out[k], ts)
            self._mark("mm kernel %d" % s, self.tail_ts)
            self.mm_send[k] = None if final else dist.isend(self.mm_out[k], dst=p.right, group=self.mm_group)
        _dbg("step %d: correlator" % s)
        with torch.cuda.stream(self.corr_ts):
            cs = self.corr_ts.cuda_stream
            if self.corr_send[k] is not None:
                self.corr_send[k].wait()
            if pre is not None:
                pre[2].wait()
                self._pre = None
            elif has_left:
                dist.recv(self.corr_in, src=p.left, group=self.corr_group)
            self._mark("corr state in %d" % s, self.corr_ts)
            ch.process_tail_corr_device(self.corr_in if has_left else None, self.corr_out[k], cs)
            self._mark("corr kernel %d" % s, self.corr_ts)
            self.corr_send[k] = None if final else dist.isend(self.corr_out[k], dst=p.right, group=self.corr_group)
        _dbg("step %d: done" % s)

    def _mark(self, label, stream):
        if _debug():
            e = torch.cuda.Event()
            e.record(stream)
            self._marks.append((label, e))

    def report(self, seconds=5.0):
        """GRB_SHARD_DEBUG: which of the marked points the device has reached `seconds` from now."""
        if not _debug():
            return
        time.sleep(seconds)
        pend = [l for l, e in self._marks if not e.query()]
        _dbg("reached %d of %d marks; pending: %s" % (len(self._marks) - len(pend), len(self._marks), pend[:12]))

    def prepost(self, next_step):
        """Call before a device-wide synchronisation that is followed by more steps.  Only rank 0's loop state comes from
        the PREVIOUS step (the last rank's block): that send is already in flight and can only complete against a posted
        receive -- a device-wide synchronisation would wait for it for ever.  Every other rank's left neighbour produces
        the state during the step itself."""
        p = self.plan
        if self.ring is not None:      # flags in memory need no posted receive
            return
        if p.rank != 0 or not p.has_left_state(next_step) or self._pre is not None:
            return
        with torch.cuda.stream(self.tail_ts):
            w1 = dist.irecv(self.mm_in, src=p.left, group=self.mm_group)
        with torch.cuda.stream(self.corr_ts):
            w2 = dist.irecv(self.corr_in, src=p.left, group=self.corr_group)
        self._pre = (next_step, w1, w2)

    def drain(self):
        """Orders the current stream after everything this rank has launched (kernels; the state sends complete on
        NCCL's own streams once the neighbour has posted its receive, see prepost)."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.tail_ts)
        cur.wait_stream(self.corr_ts)
        cur.wait_stream(self.halo_ts)
        self.ch.join(cur.cuda_stream)

    def gather_hits(self):
        """Every rank's accumulated hit list -> rank 0 (NCCL send/recv): returns an int64 [n, 2] tensor of
        (channel, absolute bit index) on rank 0, None elsewhere."""
        import numpy as np
        hits, n = self.ch.read_hits_array(self.ch.max_hits())
        assert n <= self.ch.max_hits(), "hit list overflow: %d > %d" % (n, self.ch.max_hits())
        mine = torch.empty((n, 2), dtype=torch.int64, device=self.dev)
        if n:
            mine[:, 0] = torch.from_numpy(hits["channel"].astype(np.int64)).to(self.dev)
            mine[:, 1] = torch.from_numpy(hits["bit_index"].astype(np.int64)).to(self.dev)
        p = self.plan
        sizes = [torch.zeros(1, dtype=torch.int64, device=self.dev) for _ in range(p.world)]
        dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device=self.dev))
        if p.rank == 0:
            parts = [mine]
            for r in range(1, p.world):
                t = torch.empty((int(sizes[r].item()), 2), dtype=torch.int64, device=self.dev)
                if t.numel():
                    dist.recv(t, src=r)
                parts.append(t)
            return torch.cat(parts)
        if n:
            dist.send(mine, dst=0)
        return None


def hit_checksum(t):
    """Order-independent 64-bit checksum of an [n, 2] (channel, bit index) tensor + the count."""
    if t is None or t.numel() == 0:
        return 0, 0
    key = t[:, 0] * 1000003 + t[:, 1] * 7919 + 12345
    mixed = (key ^ (key >> 13)) * 0x9E3779B1
    return int(mixed.sum().item() & 0x7FFFFFFFFFFFFFFF), int(t.shape[0])
