"""Single-process stand-in for accelerate.Accelerator (what the reference trainer calls)."""
import contextlib
import enum
import torch
from . import utils  # noqa: F401


class DistributedType(enum.Enum):
    NO = "NO"
    MULTI_GPU = "MULTI_GPU"


class _Opt:
    """accelerate's AcceleratedOptimizer exposes .scaler; None when not fp16."""

    def __init__(self, opt):
        self.optimizer = opt
        self.scaler = None

    def __getattr__(self, k):
        return getattr(self.optimizer, k)

    def zero_grad(self, *a, **k):
        return self.optimizer.zero_grad(*a, **k)

    def step(self, *a, **k):
        return self.optimizer.step(*a, **k)

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        return self.optimizer.load_state_dict(sd)


class Accelerator:
    def __init__(self, kwargs_handlers=None, mixed_precision="no", device=None, **_):
        self.mixed_precision = mixed_precision
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.is_main_process = True
        self.is_local_main_process = True
        self.distributed_type = DistributedType.NO
        self.num_processes = 1

    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, torch.optim.Optimizer):
                out.append(_Opt(o))
            elif isinstance(o, torch.nn.Module):
                out.append(o.to(self.device))
            else:
                out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    @contextlib.contextmanager
    def autocast(self):
        if self.mixed_precision in ("bf16", "fp16"):
            dt = torch.bfloat16 if self.mixed_precision == "bf16" else torch.float16
            with torch.autocast(self.device.type, dtype=dt):
                yield
        else:
            yield

    def backward(self, loss, **kw):
        loss.backward(**kw)

    def unwrap_model(self, m):
        return m

    def wait_for_everyone(self):
        pass

    def print(self, *a, **k):
        print(*a, **k)
