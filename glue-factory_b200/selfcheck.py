"""`__graft_entry__.smoke()`: one small matcher training step on cuda:0 through the public plugin
API, checked against the CPU oracle (the oracle is only the checker here)."""
import os
import sys

import torch

from . import synthetic
from .matchers.lightglue import LightGlue
from .trainer import MatcherTrainer


def smoke(verbose=True):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import lightglue_oracle as O  # checker only

    dev = torch.device("cuda:0")
    conf = dict(synthetic.DEFAULT_CONF, n_layers=2)
    weights = synthetic.make_weights(conf, seed=3)
    data = synthetic.make_pairs(2, 256, seed=5)
    with torch.no_grad():
        w64 = {k: v.double() for k, v in weights.items()}
        d64 = synthetic.to_device({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v)
                                   for k, v in data.items() if k not in ("view0", "view1")}, "cpu")
        d64["view0"] = {"image_size": data["view0"]["image_size"].double()}
        d64["view1"] = {"image_size": data["view1"]["image_size"].double()}
        ref_pred = O.lightglue_forward(w64, d64, conf)
        ref_loss = O.lightglue_loss(w64, ref_pred, d64, conf)["total"].mean().item()
    for precision, tol in (("fp32", 1e-3), ("bf16", 5e-2)):
        model = LightGlue(dict(conf, precision=precision))
        model.load_state_dict(weights, strict=False)
        model = model.to(dev)
        model.eval()
        with torch.no_grad():
            pred = model(synthetic.to_device(data, dev))
        agree = (pred["matches0"].cpu() == ref_pred["matches0"]).float().mean().item()
        # row argmax must agree wherever the oracle's top-2 margin is above the path's rounding noise
        inner = ref_pred["log_assignment"][:, :-1, :-1]
        top2 = inner.topk(2, dim=2).values
        safe = (top2[..., 0] - top2[..., 1]) > (1e-4 if precision == "fp32" else 0.3)
        got = pred["log_assignment"][:, :-1, :-1].max(2).indices.cpu()
        ok = torch.equal(got[safe], inner.max(2).indices[safe])
        if verbose:
            print(f"[smoke] precision={precision} matches0 agreement with oracle = {agree:.4f}; "
                  f"row argmax equal on {int(safe.sum())}/{safe.numel()} well-separated rows: {ok}")
        assert ok and (agree == 1.0 or precision != "fp32")
        trainer = MatcherTrainer(model, lr=1e-4)  # forward + loss + backward + gradient exchange + Adam
        loss, _ = trainer.step(data, device=dev)
        torch.cuda.synchronize()
        got = loss.item()
        err = abs(got - ref_loss) / abs(ref_loss)
        if verbose:
            print(f"[smoke] precision={precision} loss={got:.6f} oracle={ref_loss:.6f} rel.err={err:.2e}")
        assert err < tol, f"smoke: {precision} loss {got} vs oracle {ref_loss}"
    return True
