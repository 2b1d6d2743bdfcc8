"""A small reverse-mode tape over the C-ABI kernels — what tf.gradients is to the reference
(tools/train_utils.py:314,383).  Layer functions called with is_training=True append one backward
closure each; `Tape.backward()` replays them in reverse.  Gradients of activations are keyed by
tensor identity, gradients of variables accumulate into `store.grads[name]`.
"""
import contextlib

_current = None


class Tape:
    def __init__(self, store):
        self.store = store
        self.ops = []            # [(output_tensor, backward_fn(grad_out))]
        self.grads = {}          # id(tensor) -> grad tensor
        self.keep = []           # keeps tensors alive so ids stay unique

    def record(self, output, backward_fn):
        self.ops.append((output, backward_fn))
        self.keep.append(output)

    def add_grad(self, tensor, grad):
        k = id(tensor)
        if k in self.grads:
            self.grads[k] = self.grads[k] + grad
        else:
            self.grads[k] = grad
            self.keep.append(tensor)

    def needs_grad(self, tensor):
        """True if `tensor` was produced by a recorded (differentiable) op."""
        return any(o is tensor for o, _ in self.ops)

    def backward(self):
        for out, fn in reversed(self.ops):
            g = self.grads.pop(id(out), None)
            fn(g)
        self.ops.clear()
        self.grads.clear()
        self.keep.clear()


def current():
    return _current


@contextlib.contextmanager
def recording(store):
    global _current
    prev = _current
    _current = Tape(store)
    try:
        yield _current
    finally:
        _current = prev
