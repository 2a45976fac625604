"""Host-side mirror of the ``warpctc_pytorch`` call surface used by the reference
(models/pytorch_v3/ctc/ctc.py:30-69, identical in models/pytorch/ctc/ctc.py), on top of the
C ABI in include/b200ctc.h.

Surface kept (same names, argument order and meaning):
  * ``gpu_ctc(acts, grads, labels, label_lens, act_lens, minibatch_size, costs[, blank])``
        fills ``grads`` (CUDA) and ``costs`` (CPU, per utterance) in place; note the order
        label_lens BEFORE act_lens (reference call, ctc.py:39-45).
  * ``cpu_ctc(...)``   raises: the product has no CPU path (the CPU restatement lives in oracle/).
  * ``_CTC``           autograd.Function; ``forward`` may be overridden by a subclass that only
        stashes ``ctx.grads`` (exactly what the reference does, ctc.py:30-52) while ``backward``
        is inherited from here.
  * ``CTCLoss(size_average=False, length_average=False, blank=0)``  nn.Module front end.
Native additions: ``ctc_loss_and_grad`` (device-resident results, no host sync) and
``ctc_loss`` (autograd, device scalar).
"""

import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import B200CTCError

_HANDLES = {}
_WORKSPACES = {}     # (device index, stream handle) -> growing uint8 workspace tensor
_EXT = None          # the PyTorch C++ extension over the C ABI (csrc/torch_binding.cpp), False when not built


def _ext():
    """The thin PyTorch C++ extension (north_star: "a thin PyTorch C++/CUDA extension over a C-ABI"), built
    in-tree by ``__graft_entry__.build()`` / ``build.build_extension()``.  When it is not built (or
    ``B200CTC_BINDING=ctypes``) the ctypes binding calls the same C entry points."""
    global _EXT
    if _EXT is None:
        _EXT = False
        if os.environ.get("B200CTC_BINDING", "ext") != "ctypes":
            try:
                import importlib.util
                from . import build as _build
                _lib.load()                                   # libb200ctc.so first (also builds it when missing)
                path = _build.extension_path()
                if os.path.exists(path):
                    spec = importlib.util.spec_from_file_location(_build.EXT_NAME, path)
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    _EXT = mod
            except Exception:                                 # an extension built for another torch / python: fall back
                _EXT = False
    return _EXT


def binding():
    """'extension' or 'ctypes': which binding ``ctc_loss_and_grad`` goes through."""
    return "extension" if _ext() else "ctypes"


def _handle(device_index):
    """One library handle per device (the handle serialises its calls with a mutex, include/b200ctc.h);
    shared between the extension and the ctypes entry points (diagnostics, decoders)."""
    h = _HANDLES.get(device_index)
    if h is None:
        ext = _ext()
        if ext:
            hp = ctypes.c_void_p(ext.handle_address(int(device_index)))
        else:
            lib = _lib.load()
            hp = ctypes.c_void_p()
            _lib.check(lib.b200ctc_create(ctypes.byref(hp), int(device_index)), "b200ctc_create")
        h = _HANDLES[device_index] = hp
    return h


def _workspace(dev_index, stream, nbytes):
    """Workspace of the calls issued on one stream of one device: allocated once and grown on demand (no
    caching-allocator round trip per call).  Calls on one stream are ordered, so they can share it; calls on
    another stream get their own -- which is also what keeps the buffer alive and un-recycled while kernels
    of that stream still use it (no record_stream needed)."""
    key = (dev_index, stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=torch.device("cuda", dev_index))
        _WORKSPACES[key] = ws
    return ws


def release_workspaces():
    """Drop the cached workspaces (they are re-allocated on the next call)."""
    _WORKSPACES.clear()
    if _ext():
        _ext().release_workspaces()


def set_profiling(enable, device_index=None):
    """Bracket each kernel of the following ctc_loss_and_grad calls with CUDA events (bench only)."""
    if device_index is None:
        device_index = torch.cuda.current_device()
    _lib.check(_lib.load().b200ctc_set_profiling(_handle(device_index), 1 if enable else 0), "b200ctc_set_profiling")


def last_kernel_ms(device_index=None):
    """Device time (ms) of the last call's kernels: (softmax rows, lattice, cost sum)."""
    if device_index is None:
        device_index = torch.cuda.current_device()
    ms = (ctypes.c_float * 3)()
    _lib.check(_lib.load().b200ctc_get_last_kernel_ms(_handle(device_index), ms), "b200ctc_get_last_kernel_ms")
    return tuple(float(x) for x in ms)


def last_fallbacks(device_index=None, with_invalid=False):
    """(extreme_rows, range_lost): utterances of the last call that took the fp64 safe lattice;
    with_invalid=True appends the number of utterances a device-resident call rejected (cost NaN)."""
    if device_index is None:
        device_index = torch.cuda.current_device()
    c = (ctypes.c_int * 3)()
    _lib.check(_lib.load().b200ctc_get_last_fallbacks(_handle(device_index), c,
                                                      torch.cuda.current_stream().cuda_stream),
               "b200ctc_get_last_fallbacks")
    return (int(c[0]), int(c[1]), int(c[2])) if with_invalid else (int(c[0]), int(c[1]))


def plan_cache_stats(device_index=None):
    """(hits, misses) of the host-label call's plan cache on this device's handle."""
    if device_index is None:
        device_index = torch.cuda.current_device()
    h, m = ctypes.c_longlong(), ctypes.c_longlong()
    _lib.check(_lib.load().b200ctc_get_plan_cache_stats(_handle(device_index), ctypes.byref(h), ctypes.byref(m)),
               "b200ctc_get_plan_cache_stats")
    return int(h.value), int(m.value)


def _host_i32(x, name):
    """labels / lengths arrive as CPU int32 tensors in the reference (ctc.py:295-297,321);
    numpy arrays and lists are accepted too.  Returns a C-contiguous int32 numpy array."""
    if isinstance(x, np.ndarray) and x.dtype == np.int32 and x.ndim == 1 and x.flags.c_contiguous:
        return x
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            x = x.cpu()  # the warp-ctc contract keeps these on the host; tolerate device tensors
        x = x.detach().numpy()
    arr = np.ascontiguousarray(np.asarray(x), dtype=np.int32)
    if arr.ndim != 1:
        raise B200CTCError("%s must be 1-dimensional" % name)
    return arr


def _as_cpu_tensor(x):
    """labels / lengths for the extension's host-label entry point: a 1-D CPU int32 tensor (no copy for int32 numpy)."""
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.int32))


def _i32_ptr(arr):
    return arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _require_cuda(acts):
    if not isinstance(acts, torch.Tensor) or not acts.is_cuda:
        raise B200CTCError("acts must be a CUDA tensor: this engine has no CPU path "
                           "(the CPU restatement lives in oracle/ for tests only)")
    if acts.dtype != torch.float32:
        raise B200CTCError("acts must be float32, got %s" % acts.dtype)
    if acts.dim() != 3:
        raise B200CTCError("acts must be [T, B, V]")


def workspace_bytes(label_lens, act_lens, T, V):
    lib = _lib.load()
    label_lens = _host_i32(label_lens, "label_lens")
    act_lens = _host_i32(act_lens, "act_lens")
    n = ctypes.c_size_t()
    _lib.check(lib.b200ctc_get_workspace_size(_i32_ptr(label_lens), _i32_ptr(act_lens), int(T), int(V),
                                              len(act_lens), ctypes.byref(n)), "b200ctc_get_workspace_size")
    return n.value


def workspace_bound(T, V, B, max_label_len):
    """Workspace size from the shape alone (what the device-resident call uses)."""
    n = ctypes.c_size_t()
    _lib.check(_lib.load().b200ctc_get_workspace_bound(int(T), int(V), int(B), int(max_label_len), ctypes.byref(n)),
               "b200ctc_get_workspace_bound")
    return n.value


def _outputs(T, B, V, dev, grads, need_grad, costs, loss_sum):
    if need_grad:
        if grads is None:
            grads = torch.empty((T, B, V), dtype=torch.float32, device=dev)
        elif (not grads.is_cuda or grads.dtype != torch.float32 or tuple(grads.shape) != (T, B, V)
              or not grads.is_contiguous()):
            raise B200CTCError("grads must be a contiguous CUDA float32 tensor shaped like acts")
    else:
        grads = None
    if costs is None:
        costs = torch.empty(B, dtype=torch.float32, device=dev)
    if loss_sum is None:
        loss_sum = torch.empty(1, dtype=torch.float32, device=dev)
    return grads, costs, loss_sum


def _dev_i32(x, name, dev):
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise B200CTCError("%s must be a CUDA tensor when the labels are device-resident" % name)
    if x.dtype != torch.int32 or not x.is_contiguous() or x.device != dev:
        x = x.to(device=dev, dtype=torch.int32).contiguous()
    return x


def ctc_loss_and_grad(acts, labels, act_lens, label_lens, blank=0, grads=None, need_grad=True,
                      costs=None, loss_sum=None, max_label_len=None, logit_scale=1.0, label_smoothing=0.0,
                      loss_scale=1.0, grad_scale=1.0, ls_costs=None):
    """One fused cost-and-gradient evaluation on the current CUDA stream.

    acts [T,B,V] CUDA fp32 logits (any strides with a unit vocabulary stride, e.g. the
    ``logits.transpose(0, 1)`` view the reference passes, ctc.py:319 -- no copy is made).
    Two forms of labels / lengths:
      * host-resident (the warp-ctc contract, ctc.py:295-297,321): flat int32 ``labels`` and ``act_lens`` /
        ``label_lens`` as CPU tensors, numpy arrays or lists -- planned on the host, tables copied by the call;
      * device-resident: ``labels`` a CUDA int32 tensor ``[B, Lmax]`` (padded; entries past ``label_lens[b]`` are
        ignored) with CUDA int32 ``act_lens`` / ``label_lens`` -- planned by a kernel, the call is kernel
        launches only and can be captured into a CUDA graph (``max_label_len`` defaults to ``Lmax``).
    ``logit_scale`` / ``label_smoothing`` / ``loss_scale`` / ``grad_scale`` (device-resident form only) fuse the
    arithmetic of the reference's call site into the kernels (b200ctc_options, include/b200ctc.h):
    logits * logit_scale (1/temperature, ctc.py:306-307), loss_sum = loss_scale * sum_b[(1-ls) cost_b + ls/V * XE_b]
    (ctc.py:323,329-337), grads = grad_scale * d(loss_sum/loss_scale)/d(scaled logits).
    Returns device tensors ``(costs[B], loss_sum[1], grads[T,B,V] or None)``; nothing synchronises the host.
    """
    _require_cuda(acts)
    ext = _ext()
    if ext:
        try:
            if isinstance(labels, torch.Tensor) and labels.is_cuda:
                return ext.loss_and_grad_dev(acts, labels, act_lens, label_lens, int(blank), grads, bool(need_grad), costs,
                                             loss_sum, -1 if max_label_len is None else int(max_label_len),
                                             float(logit_scale), float(label_smoothing), float(loss_scale),
                                             float(grad_scale), ls_costs)
            if logit_scale != 1.0 or label_smoothing != 0.0 or loss_scale != 1.0 or grad_scale != 1.0:
                raise B200CTCError("logit_scale / label_smoothing / loss_scale / grad_scale need device-resident labels "
                                   "(the warp-ctc style call has no such arguments); see ctc_loss_from_padded")
            return ext.loss_and_grad_host(acts, _as_cpu_tensor(labels), _as_cpu_tensor(act_lens), _as_cpu_tensor(label_lens),
                                          int(blank), grads, bool(need_grad), costs, loss_sum)
        except B200CTCError:
            raise
        except (RuntimeError, TypeError) as exc:
            raise B200CTCError(str(exc).split("\n")[0])
    lib = _lib.load()
    if acts.stride(2) != 1 and acts.size(2) > 1:
        acts = acts.contiguous()
    T, B, V = acts.shape
    dev = acts.device
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    switch = torch.cuda.current_device() != dev_index
    if switch:
        prev_dev = torch.cuda.current_device()
        torch.cuda.set_device(dev_index)
    try:
        grads, costs, loss_sum = _outputs(T, B, V, dev, grads, need_grad, costs, loss_sum)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if isinstance(labels, torch.Tensor) and labels.is_cuda:
            if labels.dim() != 2 or labels.size(0) != B:
                raise B200CTCError("device-resident labels must be a padded [B, Lmax] tensor")
            labels = labels if (labels.dtype == torch.int32 and labels.stride(1) == 1) else labels.to(torch.int32).contiguous()
            act_lens_d = _dev_i32(act_lens, "act_lens", dev)
            label_lens_d = _dev_i32(label_lens, "label_lens", dev)
            if act_lens_d.numel() != B or label_lens_d.numel() != B:
                raise B200CTCError("act_lens and label_lens must have one entry per utterance (B=%d)" % B)
            Lmax = labels.size(1) if max_label_len is None else int(max_label_len)
            if Lmax > labels.size(1):
                raise B200CTCError("max_label_len exceeds the padded label width")
            nbytes = ctypes.c_size_t()
            _lib.check(lib.b200ctc_get_workspace_bound(T, V, B, Lmax, ctypes.byref(nbytes)), "b200ctc_get_workspace_bound")
            workspace = _workspace(dev_index, stream, nbytes.value)
            opts = None
            if logit_scale != 1.0 or label_smoothing != 0.0 or loss_scale != 1.0 or grad_scale != 1.0:
                opts = ctypes.byref(_lib.Options(float(logit_scale), float(label_smoothing), float(loss_scale),
                                                 float(grad_scale)))
            st = lib.b200ctc_loss_and_grad_dev(
                _handle(dev_index), acts.data_ptr(), acts.stride(0), acts.stride(1),
                grads.data_ptr() if grads is not None else None,
                labels.data_ptr(), labels.stride(0) if B > 0 and labels.size(1) > 0 else Lmax,
                label_lens_d.data_ptr(), act_lens_d.data_ptr(),
                T, V, B, Lmax, int(blank), opts, costs.data_ptr(), loss_sum.data_ptr(),
                ls_costs.data_ptr() if ls_costs is not None else None,
                workspace.data_ptr(), workspace.numel(), stream)
            _lib.check(st, "b200ctc_loss_and_grad_dev")
            return costs, loss_sum, grads
        if logit_scale != 1.0 or label_smoothing != 0.0 or loss_scale != 1.0 or grad_scale != 1.0:
            raise B200CTCError("logit_scale / label_smoothing / loss_scale / grad_scale need device-resident labels "
                               "(the warp-ctc style call has no such arguments); see ctc_loss_from_padded")
        labels = _host_i32(labels, "labels")
        act_lens = _host_i32(act_lens, "act_lens")
        label_lens = _host_i32(label_lens, "label_lens")
        if len(act_lens) != B or len(label_lens) != B:
            raise B200CTCError("act_lens and label_lens must have one entry per utterance (B=%d)" % B)
        if int(label_lens.sum()) != len(labels):
            raise B200CTCError("sum(label_lens)=%d does not match len(labels)=%d" % (int(label_lens.sum()), len(labels)))
        nbytes = ctypes.c_size_t()
        _lib.check(lib.b200ctc_get_workspace_size(_i32_ptr(label_lens), _i32_ptr(act_lens), T, V, B,
                                                  ctypes.byref(nbytes)), "b200ctc_get_workspace_size")
        workspace = _workspace(dev_index, stream, nbytes.value)
        st = lib.b200ctc_loss_and_grad(
            _handle(dev_index),
            acts.data_ptr(), acts.stride(0), acts.stride(1),
            grads.data_ptr() if grads is not None else None,
            _i32_ptr(labels), _i32_ptr(label_lens), _i32_ptr(act_lens),
            T, V, B, int(blank),
            costs.data_ptr(), loss_sum.data_ptr(),
            workspace.data_ptr(), workspace.numel(), stream)
        _lib.check(st, "b200ctc_loss_and_grad")
    finally:
        if switch:
            torch.cuda.set_device(prev_dev)
    return costs, loss_sum, grads


# ------------------------------------------------------------------------------------------------
# warpctc_pytorch surface
# ------------------------------------------------------------------------------------------------

def gpu_ctc(acts, grads, labels, label_lens, act_lens, minibatch_size, costs, blank=0):
    """Drop-in for ``warpctc_pytorch.gpu_ctc`` (reference call: ctc.py:35,39-45).
    ``grads`` (CUDA, same shape as acts) and ``costs`` (CPU float32 [B]) are filled in place."""
    _require_cuda(acts)
    if int(minibatch_size) != acts.size(1):
        raise B200CTCError("minibatch_size=%d does not match acts.size(1)=%d" % (minibatch_size, acts.size(1)))
    dev_costs, _, _ = ctc_loss_and_grad(acts, labels, act_lens, label_lens, blank=blank, grads=grads)
    costs.copy_(dev_costs)  # device -> host: the legacy surface returns per-utterance costs on the CPU
    return 0


def cpu_ctc(*_args, **_kwargs):
    raise B200CTCError("cpu_ctc: this engine is CUDA-only (sm_100a); there is no CPU fallback")


class _CTC(torch.autograd.Function):
    """``warpctc_pytorch._CTC``: cost in forward, gradient stashed on ``ctx.grads``."""

    # True reproduces the oldest upstream revision, which ignored grad_output
    legacy_ignore_grad_output = False

    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, size_average=False, length_average=False, blank=0):
        if not acts.is_cuda:
            cpu_ctc()
        minibatch_size = acts.size(1)
        costs, loss_sum, grads = ctc_loss_and_grad(acts, labels, act_lens, label_lens, blank=blank)
        loss = loss_sum.cpu()  # FloatTensor[1] on the host, as upstream returns it
        if length_average:
            total_length = float(_host_i32(act_lens, "act_lens").sum())
            grads = grads / total_length
            loss = loss / total_length
        elif size_average:
            grads = grads / minibatch_size
            loss = loss / minibatch_size
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        grads = ctx.grads
        if isinstance(grads, torch.Tensor) and not _CTC.legacy_ignore_grad_output:
            grads = grads * grad_output.to(grads.device).reshape(-1)[0]
        # one slot per forward input: 7 here, 5 for the reference's override (ctc.py:32)
        return (grads,) + (None,) * (len(ctx.needs_input_grad) - 1)


class CTCLoss(torch.nn.Module):
    """``warpctc_pytorch.CTCLoss`` (instantiated by the reference at ctc.py:69).

    size_average: divide cost and gradient by the mini-batch size;
    length_average: divide by the total number of frames; blank: blank symbol index.
    """

    def __init__(self, size_average=False, length_average=False, blank=0):
        super().__init__()
        self.ctc = _CTC.apply
        self.size_average = size_average
        self.length_average = length_average
        self.blank = blank

    def forward(self, acts, labels, act_lens, label_lens):
        if labels.dim() != 1:
            raise B200CTCError("labels must be 1 dimensional")
        for name, x in (("labels", labels), ("act_lens", act_lens), ("label_lens", label_lens)):
            if isinstance(x, torch.Tensor) and x.requires_grad:
                raise B200CTCError("%s must not require gradients" % name)
        return self.ctc(acts, labels, act_lens, label_lens, self.size_average, self.length_average, self.blank)


# ------------------------------------------------------------------------------------------------
# native API: device-resident loss, no host synchronisation
# ------------------------------------------------------------------------------------------------

class _CTCDevice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, blank, reduction):
        costs, loss_sum, grads = ctc_loss_and_grad(acts, labels, act_lens, label_lens, blank=blank,
                                                   need_grad=acts.requires_grad)
        ctx.grads = grads
        ctx.reduction = reduction
        if reduction == "none":
            return costs
        if reduction == "mean":
            return loss_sum / acts.size(1)
        return loss_sum

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.reduction == "none":
            g = ctx.grads * grad_output.reshape(1, -1, 1)
        elif ctx.reduction == "mean":
            g = ctx.grads * (grad_output.reshape(-1)[0] / ctx.grads.size(1))
        else:
            g = ctx.grads * grad_output.reshape(-1)[0]
        return g, None, None, None, None, None


def ctc_loss(acts, labels, act_lens, label_lens, blank=0, reduction="sum"):
    """Device-resident CTC loss: returns a CUDA tensor ([1] for 'sum'/'mean', [B] for 'none')."""
    if reduction not in ("sum", "mean", "none"):
        raise B200CTCError("reduction must be 'sum', 'mean' or 'none'")
    return _CTCDevice.apply(acts, labels, act_lens, label_lens, blank, reduction)


# ------------------------------------------------------------------------------------------------
# call-site helpers: what CTC.forward does around the op, without the python loop,
# the transpose copy and the host round trips
# ------------------------------------------------------------------------------------------------

def concatenate_labels(ys, y_lens):
    """Vectorised ``_concatenate_labels`` (reference: models/pytorch_v3/ctc/ctc.py:532-549, a python loop
    over the mini-batch): padded ``ys[B, Lmax]`` + ``y_lens[B]`` -> flat int32 ``[sum(y_lens)]`` on the host."""
    ys = ys.detach().cpu().numpy() if isinstance(ys, torch.Tensor) else np.asarray(ys)
    y_lens = _host_i32(y_lens, "y_lens")
    if ys.ndim != 2 or ys.shape[0] != len(y_lens):
        raise B200CTCError("ys must be [B, Lmax] with one length per row")
    if len(y_lens) and int(y_lens.max(initial=0)) > ys.shape[1]:
        raise B200CTCError("y_lens exceeds the padded label width")
    mask = np.arange(ys.shape[1])[None, :] < y_lens[:, None]
    return np.ascontiguousarray(ys[mask], dtype=np.int32)


class _CTCFromPadded(torch.autograd.Function):
    """The whole loss computation of ``CTC.forward`` as ONE device-resident call: temperature, ``/ len(xs)`` and
    the label-smoothing cross entropy are evaluated inside the kernels (b200ctc_options)."""

    @staticmethod
    def forward(ctx, logits, ys, x_lens, y_lens, inv_temperature, label_smoothing, loss_scale):
        need_grad = logits.requires_grad
        _, loss, grads = ctc_loss_and_grad(
            logits.transpose(0, 1), ys, x_lens, y_lens, blank=0, need_grad=need_grad,
            logit_scale=inv_temperature, label_smoothing=label_smoothing, loss_scale=loss_scale,
            grad_scale=loss_scale * inv_temperature)      # d loss / d (unscaled, batch-major) logits
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        g = ctx.grads * grad_output.reshape(-1)[0]        # [T, B, V]
        return g.transpose(0, 1), None, None, None, None, None, None


def _dev_padded_labels(ys, y_lens, dev, label_offset):
    """Padded ``ys[B, Lmax]`` (+ ``label_offset``) and ``y_lens[B]`` as CUDA int32 tensors (the reference keeps
    them on the host, ctc.py:295-297: they are uploaded here, a few kilobytes)."""
    ys = torch.as_tensor(ys) if not isinstance(ys, torch.Tensor) else ys
    y_lens = torch.as_tensor(y_lens) if not isinstance(y_lens, torch.Tensor) else y_lens
    if ys.dim() != 2 or ys.size(0) != y_lens.numel():
        raise B200CTCError("ys must be [B, Lmax] with one length per row")
    ys = ys.to(device=dev, dtype=torch.int32)
    if label_offset:
        ys = ys + int(label_offset)
    return ys.contiguous(), y_lens.to(device=dev, dtype=torch.int32).contiguous()


def ctc_loss_from_padded(logits, ys, x_lens, y_lens, label_offset=1, logits_temperature=1.0, average=True,
                         label_smoothing=0.0):
    """The loss computation of ``CTC.forward`` (reference: ctc.py:299-337) as one call, device resident.

    logits ``[B, T, V]`` CUDA (batch-major, as the encoder returns them; consumed through a strided view,
    no ``transpose().contiguous()`` copy), ys ``[B, Lmax]`` padded labels WITHOUT the blank offset
    (``label_offset=1`` reproduces ``ys = ys + 1``, ctc.py:300: index 0 is the blank), x_lens / y_lens ``[B]``
    (host or device).  ``logits_temperature`` is the reference's "output smoothing" (ctc.py:306-307),
    ``average=True`` its ``/ len(xs)`` (ctc.py:323), ``label_smoothing`` its ``ls_prob`` (ctc.py:329-337 with
    models/pytorch_v3/criterion.py:51-80): all three are evaluated inside the kernels, not as tensor passes.
    Returns a CUDA tensor ``[1]``:  ``(1-ls) * sum_b cost_b / B + ls/V * sum_b sum_{t<x_lens[b]} sum_k -lp[b,t,k] / B``
    that back-propagates into ``logits``.
    """
    if logits.dim() != 3:
        raise B200CTCError("logits must be [B, T, V]")
    _require_cuda(logits)
    dev = logits.device
    ys_d, y_lens_d = _dev_padded_labels(ys, y_lens, dev, label_offset)
    x_lens_d = (x_lens if isinstance(x_lens, torch.Tensor) else torch.as_tensor(np.asarray(x_lens))).to(
        device=dev, dtype=torch.int32).contiguous()
    B = logits.size(0)
    loss_scale = 1.0 / B if (average and B > 0) else 1.0
    return _CTCFromPadded.apply(logits, ys_d, x_lens_d, y_lens_d, 1.0 / float(logits_temperature),
                                float(label_smoothing), loss_scale)
