"""Time sharding of the wideband stream across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests).  The stream is cut
into contiguous time blocks of `rows_per_block` channel-rate rows; block b = step*world + rank goes to
`rank`.  Two things cross rank boundaries, both as point-to-point send/recv on their own process
group so that their message orders cannot interleave:

  halo   the last `halo_rows` INPUT rows of the left neighbour's block (tap history of the
         channelizer + enough extra rows to rebuild the discriminator / matched-filter histories and
         the M&M look-back locally).  Finite-memory stages need nothing else.
  state  the per-channel LOOP state {mu, omega, last_sample, next input index, slicer avg,
         correlator registers}: the M&M / slicer / correlator recurrences have infinite memory, so
         block b's tail stage starts from block b-1's final state.  This is a ring: rank r receives
         from r-1 and sends to r+1; rank 0 receives what rank world-1 sent in the previous step.

The front stage of every rank runs concurrently; only the (cheap, latency-bound) tail stages are
chained.  Nothing here computes: the data path is libgr_cuda.
"""
import torch
import torch.distributed as dist


class TimeShardPlan:
    def __init__(self, world, rank, rows_per_block, halo_rows):
        self.world, self.rank = int(world), int(rank)
        self.rows_per_block, self.halo_rows = int(rows_per_block), int(halo_rows)

    def block_index(self, step):
        return step * self.world + self.rank

    def abs_start(self, step):
        """Absolute channel-rate row index of the first new row of this rank's block at `step`."""
        return self.block_index(step) * self.rows_per_block

    @property
    def left(self):
        return (self.rank - 1) % self.world

    @property
    def right(self):
        return (self.rank + 1) % self.world

    @property
    def front_after_own_tail(self):
        """Scheduling policy of a rank's two streams.  The tails of all blocks form one serial chain, so a rank's tail
        occupies one slot in `world`.  Running a front underneath the rank's own tail slows both down (front 1.45 ->
        1.78 ms, tail 0.84 -> 1.24 ms on a B200); ordering the front strictly AFTER the rank's previous tail keeps both
        at their stand-alone speed, and with 2+ ranks the device still has work while it waits for the neighbour's
        state: period = max(tail + front, world x (tail + hand-off)).  A single rank has nobody to wait for and
        overlaps the two instead (1.86 ms per block against 2.37 ms in sequence)."""
        return self.world >= 2

    def has_left_state(self, step):
        """False only for the very first block of the stream (it starts from the constructor state)."""
        return self.block_index(step) > 0


class RingExchanger:
    """Halo + loop-state exchange between neighbouring time shards."""

    def __init__(self, plan, halo_group=None, state_group=None):
        self.plan = plan
        if plan.world > 1:
            self.halo_group = halo_group if halo_group is not None else dist.new_group()
            self.state_group = state_group if state_group is not None else dist.new_group()
        else:
            self.halo_group = self.state_group = None
        self._pending = None     # the state send in flight (non-blocking: see send_state)
        self._send_buf = None
        self._prerecv = None     # (step, work) of a receive posted ahead of its step

    def exchange_halo(self, my_tail_rows, halo_out, step):
        """Sends the last halo rows of my block to the right neighbour and receives my left neighbour's
        into `halo_out`.  Returns the list of pending works (call wait_all before the front stage)."""
        p = self.plan
        if p.world == 1:
            return []
        ops = [dist.P2POp(dist.isend, my_tail_rows, p.right, group=self.halo_group),
               dist.P2POp(dist.irecv, halo_out, p.left, group=self.halo_group)]
        return dist.batch_isend_irecv(ops)

    def prepost_recv(self, state_buf, step):
        """Posts the receive of the state `step` will need without waiting for it.  Call it before a device-wide
        synchronisation that has to outlive a step boundary (e.g. between warm-up and timed steps): the left
        neighbour's pending send can only complete against a posted receive."""
        p = self.plan
        # Only rank 0's state comes from the PREVIOUS step (the last rank's block); every other rank's left
        # neighbour produces it during `step` itself, so a receive posted ahead would outlive the synchronisation.
        if p.world == 1 or p.rank != 0 or not p.has_left_state(step) or self._prerecv is not None:
            return
        self._prerecv = (step, dist.irecv(state_buf, src=p.left, group=self.state_group))

    def recv_state(self, state_buf, step):
        p = self.plan
        if p.world == 1 or not p.has_left_state(step):
            return False
        if self._prerecv is not None and self._prerecv[0] == step:
            self._prerecv[1].wait()
            self._prerecv = None
            return True
        dist.recv(state_buf, src=p.left, group=self.state_group)
        return True

    def send_state(self, state_buf, step, last_step):
        p = self.plan
        if p.world == 1:
            return
        # the last block of the whole run has nobody waiting for its state
        if step == last_step and p.rank == p.world - 1:
            return
        # Non-blocking, from a private copy: the right neighbour only posts its recv after ITS halo exchange
        # of the next step, which in turn needs this rank to have reached the same exchange -- a blocking
        # send here deadlocks on transports whose send waits for the matching recv (gloo; NCCL only avoids
        # it because its sends are stream-ordered).
        self.finish()
        if self._send_buf is None or self._send_buf.shape != state_buf.shape or self._send_buf.device != state_buf.device:
            self._send_buf = torch.empty_like(state_buf)
        self._send_buf.copy_(state_buf)
        self._pending = dist.isend(self._send_buf, dst=p.right, group=self.state_group)

    def finish(self):
        """Waits for the state send in flight (call once after the last step)."""
        if self._pending is not None:
            self._pending.wait()
            self._pending = None

    @staticmethod
    def wait_all(works):
        for w in works:
            w.wait()


def gather_counts(value, world, device):
    """Final result gather: per-rank scalar (e.g. sync hits of the step) -> list on rank 0."""
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    if world == 1:
        return [int(value)]
    out = [torch.zeros_like(t) for _ in range(world)] if dist.get_rank() == 0 else None
    dist.gather(t, out, dst=0)
    return [int(o.item()) for o in out] if out is not None else None
