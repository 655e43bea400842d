"""Global gradient-recording switch (reference: DeepFlows/autograd.py:3-77).

One process-global flag; `no_grad` / `enable_grad` work both as context managers and as
decorators. `Module.train()/eval()` flips the same flag (reference quirk Q9, module.py:764)."""
import functools

_recording = True


def is_grad_enable():
    return _recording


def set_grad_enabled(mode: bool):
    global _recording
    _recording = bool(mode)


class _GradMode:
    _target = True

    def __enter__(self):
        self._saved = _recording
        set_grad_enabled(self._target)

    def __exit__(self, *exc):
        set_grad_enabled(self._saved)

    def __call__(self, fn):
        cls = type(self)

        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            with cls():
                return fn(*args, **kwargs)

        return wrapped


class no_grad(_GradMode):
    _target = False


class enable_grad(_GradMode):
    _target = True
