"""Flat gradient arena of one student + the overlapped gradient average of the data-parallel step
(reference: train_final_voc.py:155,470-471 — DistributedDataParallel's reducer buckets the gradients and all-reduces them
while the backward pass is still running).

The backward kernels write every parameter gradient straight into ONE fp32 buffer per student, laid out in the order in
which the backward pass finishes them (heads and decoder first, encoder blocks 11 .. 0, patch embedding last), and the
parameters' `.grad` are views of it.  As soon as a run of consecutive gradients of at least `chunk_elems` elements is
final, its slice of the arena is handed to an asynchronous all-reduce (mean over the ranks, NCCL on its own stream), so
NVLink time hides behind the remaining wgrad / dgrad GEMMs; nothing is flattened or copied back.  With one rank the
arena only removes the per-tensor allocations.  Works inside CUDA-graph capture (the NCCL stream forks off the capturing
stream through events and is joined by `finish()`).
"""
import torch


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class GradArena:
    ALIGN = 64  # elements: every view starts on a 256-byte boundary (128-bit stores of the GEMM epilogue, NCCL alignment)

    def __init__(self, named_params, order, chunk_elems=6 << 20, device=None, group=None):
        """named_params: [(name, parameter)] of the trainable parameters; order: the same names in the order the backward
        pass completes their gradients."""
        params = dict(named_params)
        if set(order) != set(params):
            raise ValueError("backward order does not cover the trainable parameters: "
                             f"{sorted(set(order) ^ set(params))[:4]}")
        self.names = list(order)
        self.params = params
        self.group = group
        dev = device or next(iter(params.values())).device
        self.offsets, off = {}, 0
        for n in self.names:
            self.offsets[n] = off
            off += (params[n].numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.views = {n: self.flat[self.offsets[n]:self.offsets[n] + params[n].numel()].view(params[n].shape) for n in self.names}
        # chunks: runs of consecutive names with >= chunk_elems elements (the last one takes the remainder)
        self.chunks, cur, lo = [], [], 0
        for i, n in enumerate(self.names):
            cur.append(n)
            hi = self.offsets[self.names[i + 1]] if i + 1 < len(self.names) else self.total
            if hi - lo >= chunk_elems or i + 1 == len(self.names):
                self.chunks.append((lo, hi, tuple(cur)))
                cur, lo = [], hi
        self._chunk_of = {n: ci for ci, (_, _, ns) in enumerate(self.chunks) for n in ns}
        self.ever_written = set()
        self.begin_step(1)

    # ------------------------------------------------------------------ per-step protocol
    def begin_step(self, expected_backwards=1):
        """Call before the backward pass(es) of a step: `expected_backwards` StudentFunction.backward calls will hit this
        arena (2 in phase C: the plain and the augmented view).  The first one writes, later ones accumulate; the all-reduce
        of a chunk is issued during the last one."""
        self.expected = expected_backwards
        self.calls_done = 0
        self._pending = [len(ns) for (_, _, ns) in self.chunks]
        self._works = []
        self._issued = 0

    @property
    def writing(self):
        return self.calls_done == 0

    @property
    def last_call(self):
        return self.calls_done + 1 >= self.expected

    def out(self, name):
        """Buffer a kernel may write the gradient of `name` into directly (first backward of the step), else None."""
        return self.views[name] if self.writing else None

    def put(self, name, g):
        """The gradient of `name` for this backward call is final: `g` is either the arena view itself (written in place) or
        a tensor to copy / accumulate.  g=None: this call contributes nothing (the view is zeroed if this is the writing call)."""
        v = self.views[name]
        if g is None:
            if self.writing:
                v.zero_()
        else:
            self.ever_written.add(name)
            if g.data_ptr() != v.data_ptr():
                g = g.reshape(v.shape)
                if self.writing:
                    v.copy_(g)
                else:
                    v.add_(g)
            elif not self.writing:
                raise RuntimeError("in-place gradient write during an accumulating backward call")
        if self.last_call:
            ci = self._chunk_of[name]
            self._pending[ci] -= 1
            if self._pending[ci] == 0:
                self._flush_ready()

    def _flush_ready(self):
        # chunks are issued strictly in arena order so that every rank enqueues the same sequence of collectives
        while self._issued < len(self.chunks) and self._pending[self._issued] == 0:
            lo, hi, _ = self.chunks[self._issued]
            self._issued += 1
            self._reduce(lo, hi)

    def _reduce(self, lo, hi):
        if _world() == 1:
            return
        import torch.distributed as dist
        t = self.flat[lo:hi]
        if dist.get_backend(self.group) == "nccl":
            self._works.append((dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True), None))
        else:   # gloo (CPU tests) has no AVG
            self._works.append((dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True), t))

    def end_call(self):
        """End of one StudentFunction.backward: gradients this call never produced count as final."""
        self.calls_done += 1

    def finish(self):
        """After loss.backward(): issue whatever is left (parameters without a gradient this step leave their chunk open) and
        make the current stream wait for every outstanding all-reduce."""
        if self.calls_done >= self.expected or self.calls_done == 0:
            for ci in range(len(self.chunks)):
                self._pending[ci] = 0
            self._flush_ready()
        world = _world()
        for work, scale in self._works:
            work.wait()
            if scale is not None:
                scale.div_(world)
        self._works = []

    def bind_grads(self):
        """p.grad = view for every parameter that has received a gradient so far, None for the others (AdamW must skip
        parameters the loss does not reach in this phase, exactly like the reference: no weight decay, no step count)."""
        for n, p in self.params.items():
            p.grad = self.views[n] if n in self.ever_written else None
