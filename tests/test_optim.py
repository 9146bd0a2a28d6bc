"""Optimizer numerics: the wrapper's update math vs torch.optim and vs closed form.

The reference ships no optimizer tests at all; these pin ``SGD.optim_step`` /
``Adam.optim_step`` (``/root/reference/ps.py:197-261``) and the ``step() -> (loss, data)``
contract (``ps.py:193``).
"""
import math

import pytest
import torch

import pytorch_ps_mpi_b200 as ps
from pytorch_ps_mpi_b200.models import mnist_mlp


def _train(opt_factory, steps=4, seed=0):
    torch.manual_seed(seed)
    model = mnist_mlp(hidden=16)
    opt = opt_factory(model)
    for s in range(steps):
        g = torch.Generator().manual_seed(s)
        x, y = torch.randn(8, 784, generator=g), torch.randint(0, 10, (8,), generator=g)
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
    return [p.detach().clone() for p in model.parameters()], opt


@pytest.mark.parametrize("hyper", [
    dict(lr=0.1), dict(lr=0.05, momentum=0.9), dict(lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3),
    dict(lr=0.05, momentum=0.8, dampening=0.1, weight_decay=1e-2),
])
@pytest.mark.parametrize("mode", ["ps", "allgather"])
def test_sgd_matches_torch(hyper, mode):
    a, opt = _train(lambda m: ps.SGD(m.named_parameters(), m.parameters(), code=ps.Identity(), mode=mode, **hyper))
    b, _ = _train(lambda m: torch.optim.SGD(m.parameters(), **hyper))
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-7)
    if hyper.get("momentum"):
        assert all("momentum_buffer" in opt.state[p] for g in opt.param_groups for p in g["params"])
    opt.close()


def _ref_adam(p, g, state, lr, betas, eps, wd, amsgrad):
    """Closed form of /root/reference/ps.py:218-261 (note: sqrt(v)+eps, bias correction in step size)."""
    b1, b2 = betas
    state["t"] += 1
    if wd:
        g = g + wd * p
    state["m"] = b1 * state["m"] + (1 - b1) * g
    state["v"] = b2 * state["v"] + (1 - b2) * g * g
    v = state["v"]
    if amsgrad:
        state["vmax"] = torch.maximum(state["vmax"], v)
        v = state["vmax"]
    step = lr * math.sqrt(1 - b2 ** state["t"]) / (1 - b1 ** state["t"])
    return p - step * state["m"] / (v.sqrt() + eps)


@pytest.mark.parametrize("amsgrad", [False, True])
@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_adam_reference_math(amsgrad, wd):
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(5, 3))
    opt = ps.Adam([("w", p)], [p], lr=1e-2, betas=(0.8, 0.9), eps=1e-6, weight_decay=wd, amsgrad=amsgrad)
    assert opt.optim == "adam"            # no need to pass optim='adam' (fix of ps.py:181-188)
    want = p.detach().clone()
    st = {"t": 0, "m": torch.zeros_like(want), "v": torch.zeros_like(want), "vmax": torch.zeros_like(want)}
    for s in range(5):
        opt.zero_grad()
        (p * torch.randn(5, 3, generator=torch.Generator().manual_seed(s))).sum().backward()
        want = _ref_adam(want, p.grad.clone(), st, 1e-2, (0.8, 0.9), 1e-6, wd, amsgrad)
        opt.step()
        assert torch.allclose(p.detach(), want, rtol=1e-5, atol=1e-6)
    assert ("max_exp_avg_sq" in opt.state[p]) == amsgrad
    opt.close()


def test_adam_close_to_torch():
    a, o = _train(lambda m: ps.Adam(m.named_parameters(), m.parameters(), lr=1e-3, eps=1e-12))
    b, _ = _train(lambda m: torch.optim.Adam(m.parameters(), lr=1e-3, eps=1e-12))
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=1e-4, atol=1e-6)
    o.close()


def test_step_contract_and_closure():
    torch.manual_seed(0)
    model = mnist_mlp(hidden=8)
    opt = ps.SGD(model.named_parameters(), model.parameters(), lr=0.1, code=ps.Identity())
    x, y = torch.randn(4, 784), torch.randint(0, 10, (4,))

    def closure():
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        return loss

    loss, data = opt.step(closure)
    assert loss is not None and float(loss) > 0
    for k in ("comm_wait", "optim_step_time", "decode_time", "msg_bytes", "packaged_bytes", "code_wait",
              "iallgather_prepare_time", "isend_time"):
        assert k in data
    assert data["packaged_bytes"] >= data["msg_bytes"]          # level 0 = framing only (mpi_comms.py:18)
    assert opt.steps == 1 and opt.rank == 0 and opt.size == 1
    assert all(getattr(p, "ps_name", None) for p in model.parameters())    # ps.py:64
    assert ps.find_param(model.parameters(), "fc1.weight") is model.fc1.weight
    opt.close()


def test_sum_not_mean_and_average_flag():
    p = torch.nn.Parameter(torch.ones(4))
    opt = ps.SGD([("p", p)], [p], lr=1.0, average=True)
    p.sum().backward()
    opt.step()
    assert torch.allclose(p.detach(), torch.zeros(4))          # single rank: average == sum
    opt.close()


def test_param_groups_and_unused_params():
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(3))
    b = torch.nn.Parameter(torch.randn(3))
    unused = torch.nn.Parameter(torch.randn(2))
    opt = ps.SGD([("a", a), ("b", b), ("u", unused)],
                 [{"params": [a], "lr": 0.1}, {"params": [b, unused], "lr": 0.5}], lr=0.01)
    a0, b0, u0 = a.detach().clone(), b.detach().clone(), unused.detach().clone()
    (a.sum() + 2 * b.sum()).backward()
    opt.step()
    assert torch.allclose(a.detach(), a0 - 0.1) and torch.allclose(b.detach(), b0 - 1.0)
    assert torch.equal(unused.detach(), u0)                    # p.grad is None → skipped (ps.py:178-179)
    opt.close()


def test_duplicate_names_rejected():
    a = torch.nn.Parameter(torch.randn(3))
    b = torch.nn.Parameter(torch.randn(3))
    with pytest.raises(ValueError):
        ps.SGD([("x", a), ("x", b)], [a, b], lr=0.1)


def test_state_dict_roundtrip():
    a, opt = _train(lambda m: ps.SGD(m.named_parameters(), m.parameters(), lr=0.05, momentum=0.9))
    sd = opt.state_dict()
    assert sd["state"] and "momentum_buffer" in next(iter(sd["state"].values()))
    torch.manual_seed(0)
    m2 = mnist_mlp(hidden=16)
    o2 = ps.SGD(m2.named_parameters(), m2.parameters(), lr=0.05, momentum=0.9)
    o2.load_state_dict(sd)
    for (k1, v1), (k2, v2) in zip(sorted(sd["state"].items()), sorted(o2.state_dict()["state"].items())):
        assert torch.equal(v1["momentum_buffer"], v2["momentum_buffer"])
    opt.close(), o2.close()


def test_user_coding_slow_path():
    class Halve:
        codes = None

        def encode(self, grad, **kw):
            return {"half": grad * 0.5, "note": "user coding"}

        def decode(self, code, cuda=False):
            assert self.codes is not None and len(self.codes) == 1     # ps.py:165
            return torch.as_tensor(code["half"]) * 2

    p = torch.nn.Parameter(torch.ones(4))
    opt = ps.SGD([("p", p)], [p], lr=1.0, code=Halve(), mode="allgather")
    (3 * p).sum().backward()
    opt.step()
    assert torch.allclose(p.detach(), torch.full((4,), -2.0))
    opt.close()


def test_lr_scheduler_is_honoured():
    def run(factory):
        torch.manual_seed(0)
        model = mnist_mlp(hidden=8)
        opt = factory(model)
        sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
        for s in range(5):
            g = torch.Generator().manual_seed(s)
            x, y = torch.randn(4, 784, generator=g), torch.randint(0, 10, (4,), generator=g)
            opt.zero_grad()
            torch.nn.functional.cross_entropy(model(x), y).backward()
            opt.step()
            sched.step()
        return [p.detach().clone() for p in model.parameters()], opt

    a, o = run(lambda m: ps.SGD(m.named_parameters(), m.parameters(), lr=0.2, momentum=0.9))
    b, _ = run(lambda m: torch.optim.SGD(m.parameters(), lr=0.2, momentum=0.9))
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-7)
    assert abs(o.param_groups[0]["lr"] - 0.2 * 0.25) < 1e-12
    o.close()


def test_timings_summary_and_chrome_trace(tmp_path):
    import json
    from pytorch_ps_mpi_b200.utils import summarize_timings, dump_chrome_trace
    _, opt = _train(lambda m: ps.SGD(m.named_parameters(), m.parameters(), lr=0.1), steps=3)
    assert len(opt.timings) == 3                                  # the reference declared `timings` but never filled it (ps.py:80)
    s = summarize_timings(opt.timings)
    assert s["optim_step_time"]["n"] == 3 and s["msg_bytes"]["mean"] > 0 and s["comm_wait"]["max"] >= 0
    out = tmp_path / "trace.json"
    dump_chrome_trace(opt.timings, str(out))
    ev = json.load(open(out))["traceEvents"]
    assert ev and all(e["ph"] == "X" and e["dur"] > 0 for e in ev)
    opt.close()


def test_topk_error_feedback_through_host_optimizer():
    """the host-engine hook must hand the parameter NAME to encode() so TopK(error_feedback=True) keeps one
    residual per parameter (id(grad) of a temporary is recycled across parameters: shape errors / mixed residuals)."""
    torch.manual_seed(0)
    model = mnist_mlp(hidden=64)
    code = ps.TopK(ratio=0.05, error_feedback=True)
    opt = ps.SGD(model.named_parameters(), model.parameters(), lr=0.1, code=code, engine="host", mode="allgather")
    x, y = torch.randn(16, 1, 28, 28), torch.randint(0, 10, (16,))
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    names = {n for n, _ in model.named_parameters()}
    assert set(code._residual) == names                      # one residual per parameter, keyed by name, no growth
    for n, p in model.named_parameters():
        assert code._residual[n].numel() == p.numel()
    opt.close()


def test_user_encode_without_name_kwarg_still_works():
    class Bare(ps.Coding):
        def encode(self, grad):                              # no **kwargs: must not be handed name=
            return {"g": grad.clone()}

        def decode(self, code, cuda=False):
            return torch.as_tensor(code["g"])

    model = mnist_mlp(hidden=16)
    opt = ps.SGD(model.named_parameters(), model.parameters(), lr=0.1, code=Bare(), engine="host")
    torch.nn.functional.cross_entropy(model(torch.randn(4, 1, 28, 28)), torch.randint(0, 10, (4,))).backward()
    opt.step()
    opt.close()
