"""Batch sharding of independent simulations: over the GPUs of one box and, on one GPU, over
concurrent CUDA streams (`SampleGroups`).

The PISO path has no cross-sample coupling, so the only multi-GPU logic is: which samples does a rank own, and how are
device timings / iteration statistics combined.  No data-path collective exists; the closure-network gradient
all-reduce of the training configs belongs to the training driver (torch.distributed all_reduce over NCCL).

The same partition applies inside one GPU.  On the small grids both solvers are latency-bound persistent kernels (one
thread-block cluster per sample in the pressure CG, one CTA per system in the predictor) whose launches end in a tail of
a few slow samples while most SMs idle, and a predictor launch leaves 80 % of the issue slots unused.  `SampleGroups` runs
contiguous sample groups of the batch as independent pipelines (forward, loss, adjoint, next step ...) on their own
streams, each optionally captured once as a CUDA graph, so that the solver launches of one group fill what another
group's launch leaves idle and the host launches a handful of graphs per step instead of ~100 kernels per group.
Per-sample arithmetic is untouched: results are bit-identical to the single-stream batch (tests/test_gpu_groups.py).
Measured on B200 (periodic 128^2, batch 64, forward + adjoint): profiles/r02_stream_groups.md."""
import torch
import torch.distributed as dist


def shard_bounds(total, world_size, rank):
    """Contiguous split of `total` samples; the first `total % world_size` ranks take one extra sample."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(total), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def local_batch(tensor, world_size=None, rank=None):
    """Slice [B, ...] down to the samples this rank owns."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    start, count = shard_bounds(tensor.shape[0], world_size, rank)
    return tensor[start:start + count]


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (device timings are reported as the slowest rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def gather_counts(local_value, device=None):
    """All ranks' integer values (e.g. samples processed) as a list, for whole-job aggregates."""
    t = torch.tensor([int(local_value)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return [int(o.item()) for o in out]
    return [int(local_value)]


class SampleGroups(object):
    """Independent pipelines over contiguous sample groups of a batch, one CUDA stream (and optionally one CUDA graph)
    per group.

        runner = SampleGroups(fn, (vel, pres, w_u, w_p), groups=4, graph=True)
        for k in range(steps):
            runner.step(feedback={0: 0, 1: 1})     # outputs 0, 1 of step k are inputs 0, 1 of step k + 1
        runner.join()
        vel_next = runner.gather(0)

    `fn(*group_inputs) -> tuple of tensors` must treat the leading dimension as independent samples (a PISO step, its
    loss and its adjoint do) and must not synchronise with the host.  `inputs` are the whole-batch tensors [B, ...]
    (device or pinned host memory); each group keeps its block of every input in a static device buffer:
        load(i, *tensors)    copy new values (device or host tensors of the group's shape, None = keep) into group i's
                             input buffers, asynchronously on the group's stream
        launch(i)            run fn for group i (graph replay, or an eager call on the group's stream)
        outputs(i)           group i's output tensors (static buffers in graph mode: valid until the next launch(i))
        fetch(i, *dst)       asynchronous copy of group i's outputs into dst (e.g. pinned host buffers; None = skip)
        sync(i) / join()     the host waits for group i / the caller's current stream waits for every group
    With graph=True fn is run once eagerly (warm-up: tables, scratch, allocator) and then captured on the group's
    stream; every later launch is one cudaGraphLaunch.  Data-dependent iteration counts live inside the solver kernels,
    so the captured graph is valid for any input values."""

    def __init__(self, fn, inputs, groups=4, graph=True, device=None):
        inputs = tuple(inputs)
        if not inputs:
            raise ValueError("SampleGroups needs at least one batch tensor")
        b = int(inputs[0].shape[0])
        if any(int(t.shape[0]) != b for t in inputs):
            raise ValueError("every input must have the batch as its leading dimension")
        if device is None:
            device = next((t.device for t in inputs if t.is_cuda), None)
            if device is None:
                if not torch.cuda.is_available():
                    raise ValueError("SampleGroups runs on a CUDA device: there is no CPU path")
                device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("SampleGroups runs on a CUDA device: there is no CPU path")
        self.fn = fn
        self.groups = max(1, min(int(groups), b))
        self.bounds = [shard_bounds(b, self.groups, i) for i in range(self.groups)]
        self.graphed = bool(graph)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.groups)]
        self._in, self._out, self._graphs = [], [None] * self.groups, [None] * self.groups
        self._launched = [False] * self.groups
        caller = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(caller)
        for i, (start, count) in enumerate(self.bounds):
            st = self.streams[i]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                self._in.append([torch.empty(t[start:start + count].shape, dtype=t.dtype, device=self.device)
                                 for t in inputs])
                for buf, t in zip(self._in[i], inputs):
                    buf.copy_(t[start:start + count], non_blocking=True)
        if self.graphed:
            for i, st in enumerate(self.streams):
                with torch.cuda.stream(st):
                    self.fn(*self._in[i])                      # warm-up on the capture stream
            torch.cuda.synchronize(self.device)
            for i, st in enumerate(self.streams):
                g = torch.cuda.CUDAGraph()
                # thread_local: only the capturing thread is held to the capture rules -- a process that also runs NCCL has a
                # watchdog thread polling events, which the default "global" mode would turn into a capture error
                with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
                    out = self.fn(*self._in[i])
                self._graphs[i] = g
                self._out[i] = self._as_tuple(out)
            torch.cuda.synchronize(self.device)

    @staticmethod
    def _as_tuple(out):
        return tuple(out) if isinstance(out, (tuple, list)) else (out,)

    def inputs(self, i):
        return tuple(self._in[i])

    def outputs(self, i):
        if not self._launched[i]:
            raise RuntimeError("group %d has not been launched yet" % i)
        return self._out[i]

    def load(self, i, *tensors):
        with torch.cuda.stream(self.streams[i]):
            for buf, t in zip(self._in[i], tensors):
                if t is not None and t.data_ptr() != buf.data_ptr():
                    buf.copy_(t, non_blocking=True)

    def launch(self, i):
        with torch.cuda.stream(self.streams[i]):
            if self.graphed:
                self._graphs[i].replay()
            else:
                self._out[i] = self._as_tuple(self.fn(*self._in[i]))
        self._launched[i] = True

    def fetch(self, i, *dst):
        with torch.cuda.stream(self.streams[i]):
            for d, o in zip(dst, self.outputs(i)):
                if d is not None:
                    d.copy_(o, non_blocking=True)

    def step(self, feedback=None):
        """One launch of every group; feedback {output index: input index} first copies the previous launch's outputs
        into the inputs (a rollout: the new state is the next step's state)."""
        for i in range(self.groups):
            if feedback and self._launched[i]:
                vals = [None] * len(self._in[i])
                for o_idx, i_idx in feedback.items():
                    vals[i_idx] = self._out[i][o_idx]
                self.load(i, *vals)
            self.launch(i)

    def sync(self, i):
        self.streams[i].synchronize()

    def join(self):
        caller = torch.cuda.current_stream(self.device)
        for st in self.streams:
            e = torch.cuda.Event()
            e.record(st)
            caller.wait_event(e)

    def gather(self, k):
        """Output k of every group, concatenated along the batch on the caller's stream (joins first)."""
        self.join()
        return torch.cat([self.outputs(i)[k] for i in range(self.groups)])
