"""Host -> device staging of one batch of loss inputs (the step in front of ``Loss.forward``).

The reference uploads a batch key by key (``inputs[key] = ipt.to(self.device)``, trainer.py:226-227): ~30 small
``cudaMemcpyAsync`` calls per step on the compute stream.  ``BatchStager`` keeps the whole batch in ONE pinned host
slab and ONE device slab per buffer: the loader writes into the pinned views, ``upload()`` issues a single copy on a
dedicated copy stream, and with two (or more) buffers the copy of batch i+1 overlaps the loss kernels of batch i.
The tensors handed to the loss are views of the device slab, 256-byte aligned, contiguous fp32 NCHW as the C ABI wants.
"""
from __future__ import annotations

import contextlib
import os

import torch


def gpu_local_cpus(device):
    """The CPUs NVML lists as local to `device` (its NUMA node), or None when that cannot be determined."""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(props.uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)).encode())
        n_cpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


@contextlib.contextmanager
def numa_local(device):
    """Runs the block with the calling thread bound to the CPUs local to `device`: pinned host memory allocated inside is
    placed on the GPU's own NUMA node (the kernel's default policy is local allocation).  On a two-socket host an unbound
    process gets its pinned slab from either node, and a remote slab uploads at half the rate or less (measured on this
    pool: 14 ... 55 GB/s for the same cudaMemcpyAsync from one allocation to the next); with one rank per GPU, local slabs
    are also what lets eight uploads run at once.  No-op when NVML or the affinity calls are unavailable."""
    cpus = gpu_local_cpus(device) if hasattr(os, "sched_setaffinity") else None
    if not cpus:
        yield False
        return
    before = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
        yield True
    finally:
        os.sched_setaffinity(0, before)


def pinned_empty(nbytes, device):
    """nbytes of pinned host memory on the NUMA node of `device`, touched by a local CPU before it is handed out."""
    with numa_local(device):
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        t.zero_()
    return t


def _layout(template):
    """template: list of dicts {key: tensor}; -> ([(dict index, key, shape, dtype, offset, nbytes)], total bytes)."""
    entries, off = [], 0
    for di, d in enumerate(template):
        for k, v in d.items():
            nbytes = v.numel() * v.element_size()
            entries.append((di, k, tuple(v.shape), v.dtype, off, nbytes))
            off += (nbytes + 255) & ~255
    return entries, off


class BatchStager:
    """Double-buffered single-copy upload of a batch given as a list of ``{key: tensor}`` dicts.

    >>> st = BatchStager([inputs, flows, mobiles, cams], device)      # shapes / dtypes are taken from the template
    >>> host = st.host_views(0)                                       # pinned views the loader fills in place
    >>> dev_dicts = st.upload(0)                                      # one cudaMemcpyAsync on the copy stream
    >>> st.wait(0)                                                    # the current stream waits for that copy
    """

    def __init__(self, template, device, n_buffers=2):
        self.device = torch.device(device)
        self.entries, self.nbytes = _layout(template)
        self.n_dicts = len(template)
        self.n_buffers = n_buffers
        self.copy_stream = torch.cuda.Stream(self.device)
        self.host = [pinned_empty(self.nbytes, self.device) for _ in range(n_buffers)]     # on the GPU's own NUMA node
        self.dev = [torch.empty(self.nbytes, dtype=torch.uint8, device=self.device) for _ in range(n_buffers)]
        self.ready = [torch.cuda.Event() for _ in range(n_buffers)]     # copy of buffer k has landed
        self.free = [torch.cuda.Event() for _ in range(n_buffers)]      # consumers of buffer k have finished
        self._host_views = [self._views(h) for h in self.host]
        self._dev_views = [self._views(d) for d in self.dev]
        self._copy_pending = [False] * n_buffers   # an upload of buffer k was enqueued and not yet known to have landed
        self._in_use = [False] * n_buffers         # device buffer k was uploaded and its consumers not yet release()d
        for e in self.free:
            e.record(torch.cuda.current_stream(self.device))

    def _views(self, slab):
        out = [dict() for _ in range(self.n_dicts)]
        for di, k, shape, dtype, off, nbytes in self.entries:
            out[di][k] = slab[off:off + nbytes].view(dtype).view(shape)
        return out

    def acquire_host(self, k):
        """Blocks the calling host thread until the last enqueued copy OUT of pinned buffer k has landed: only then may
        the loader overwrite it (a non_blocking copy that is still in flight would otherwise read a torn batch)."""
        k %= self.n_buffers
        if self._copy_pending[k]:
            self.ready[k].synchronize()
            self._copy_pending[k] = False

    def host_views(self, k):
        """The pinned host tensors of buffer k (same keys / shapes as the template): write the next batch here.
        Waits (host side) for an in-flight upload of that buffer first, see acquire_host()."""
        self.acquire_host(k)
        return self._host_views[k % self.n_buffers]

    def fill(self, k, dicts):
        """Convenience for loaders that already hold tensors: copies them into the pinned views of buffer k."""
        for dst, src in zip(self.host_views(k), dicts):
            for key, v in src.items():
                dst[key].copy_(v)

    def upload(self, k):
        """Enqueues the copy of buffer k on the copy stream (after the previous consumers of that device buffer are
        done) and returns the device views.  Call wait(k) on the consuming stream before using them."""
        k %= self.n_buffers
        if self._in_use[k]:
            raise RuntimeError("BatchStager: device buffer %d is uploaded again before release(%d) marked the end of its "
                               "consumers (protocol: upload -> wait -> [kernels] -> release)" % (k, k))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[k])
            self.dev[k].copy_(self.host[k], non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self._copy_pending[k] = True
        self._in_use[k] = True
        return self._dev_views[k]

    def wait(self, k):
        torch.cuda.current_stream(self.device).wait_event(self.ready[k % self.n_buffers])

    def release(self, k):
        """Marks the end of the work that reads device buffer k (recorded on the current stream)."""
        k %= self.n_buffers
        self.free[k].record(torch.cuda.current_stream(self.device))
        self._in_use[k] = False
