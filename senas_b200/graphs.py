"""CUDA-graph capture of a whole search step (arch step + weight step, experiments/search_arc.py:252-293).

The supernet has static shapes and the step issues ~9 000 small kernels plus ~3 400 parameter tensors worth of
autograd / optimizer bookkeeping, so an eagerly launched step is bound by the host (measured on B200: enqueue time ==
step time).  Capturing forward, loss, backward, gradient clipping and both optimizer steps into one CUDA graph makes
the step cost what the kernels cost.  libsenas_b200 is capture-safe: it allocates nothing, never synchronises, and all
its plans / scratch buffers are created during the warm-up iterations that precede the capture.
"""
import torch


class GraphedSearchStep:
    """``step = GraphedSearchStep(model, criterion, w_opt, a_opt, example_batches, grad_clip=5)`` then
    ``loss = step(x_train, y_train, x_valid, y_valid)`` (device or pinned-host tensors of the captured shapes)."""

    def __init__(self, model, criterion, w_opt, a_opt, example, grad_clip=5.0, warmup=3, post_backward=None):
        xt, yt, xv, yv = example
        self.static = [t.clone() for t in (xt, yt, xv, yv)]
        self.model, self.criterion, self.w_opt, self.a_opt = model, criterion, w_opt, a_opt
        self.grad_clip, self.post_backward = grad_clip, post_backward
        for group in a_opt.param_groups:          # Adam keeps `step` on the device when capturable
            group['capturable'] = True
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step()

    def _eager_step(self):
        xt, yt, xv, yv = self.static
        self.a_opt.zero_grad(set_to_none=True)
        self.criterion(self.model(xv), yv).backward()
        if self.post_backward:
            self.post_backward()
        self.a_opt.step()
        self.w_opt.zero_grad(set_to_none=True)
        loss = self.criterion(self.model(xt), yt)
        loss.backward()
        if self.post_backward:
            self.post_backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clip)
        self.w_opt.step()
        return loss.detach()

    def __call__(self, xt, yt, xv, yv):
        for dst, src in zip(self.static, (xt, yt, xv, yv)):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
