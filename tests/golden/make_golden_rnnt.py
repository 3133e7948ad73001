"""Golden vectors for the Transducer joint + RNN-T loss, generated in the authoring container:
the joint logits come from the REAL reference module (models/joint_networks.py JointNetwork, imported from /root/reference), the loss
values from torchaudio.functional.rnnt_loss (independent implementation of the published definition the reference's warp_rnnt call
computes; warp_rnnt itself is not installable here).    python tests/golden/make_golden_rnnt.py"""
import os
import sys

import torch

REF = os.environ.get("EFFCONF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from models.joint_networks import JointNetwork          # the reference's own module (pure torch)
    import torchaudio.functional as AF
    cases = {}
    for name, (B, T, U, Denc, Ddec, J, V) in {"small": (3, 21, 7, 48, 40, 64, 37), "mid": (2, 16, 9, 168, 96, 160, 264)}.items():
        torch.manual_seed(11)
        jn = JointNetwork(Denc, Ddec, V, {"joint_mode": "sum", "dim_model": J, "act": "tanh"}).eval()
        g = torch.Generator().manual_seed(12)
        f = torch.randn(B, T, Denc, generator=g)
        gdec = torch.randn(B, U + 1, Ddec, generator=g)
        y = torch.randint(1, V, (B, U), generator=g)
        f_len = torch.tensor([T] + [max(2, T - 3 * b) for b in range(1, B)])
        y_len = torch.tensor([U] + [max(1, U - 2 * b) for b in range(1, B)])
        with torch.no_grad():
            logits = jn(f, gdec)                              # (B, T, U+1, V): 3-D inputs take the training / eval-loss branch
        per = AF.rnnt_loss(logits, y.int(), f_len.int(), y_len.int(), blank=0, reduction="none", fused_log_softmax=True)
        cases[name] = {"dims": (B, T, U, Denc, Ddec, J, V), "state_dict": {"joint_network." + k: v.clone() for k, v in jn.state_dict().items()},
                       "f": f, "g": gdec, "y": y, "f_len": f_len, "y_len": y_len, "logits": logits if name == "small" else logits[:, :4].clone(),
                       "loss_per_utt": per, "loss_mean": per.mean(),
                       "source": "logits: reference models/joint_networks.py JointNetwork; loss: torchaudio.functional.rnnt_loss " + __import__("torchaudio").__version__}
    torch.save(cases, os.path.join(HERE, "rnnt_joint_small.pt"))
    print({k: (v["dims"], float(v["loss_mean"])) for k, v in cases.items()})


if __name__ == "__main__":
    main()
