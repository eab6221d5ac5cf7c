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


class ShardedChain:
    def __init__(self, ch, plan, device, rows, halo, history_rows):
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
