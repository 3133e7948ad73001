"""CPU oracle for the Transducer joint network + RNN-T loss  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in plain torch-CPU tensor algebra (fp32 / fp64):
  * the reference's JointNetwork.forward in its training form (reference models/joint_networks.py:80-105, the shipped `joint_mode: sum`,
    `act: tanh` configuration): f (B,T,Denc), g (B,U+1,Ddec) -> Linear_enc(f)[:, :, None] + Linear_dec(g)[:, None] -> tanh -> Linear_joint
    -> logits (B, T, U+1, V);
  * the loss the reference obtains from the third-party `warp_rnnt.rnnt_loss(log_softmax(logits), labels, frames_lengths, labels_lengths,
    average_frames=False, reduction='mean', blank=0, gather=True)` (reference models/losses.py:22-46).  warp_rnnt is NOT vendored under
    /root/reference and not installed (requirements.txt lists it unpinned), so its published definition is restated: the negative log
    likelihood of Graves, "Sequence Transduction with Recurrent Neural Networks" (2012), eq. 16-18:
        alpha(0,0) = 0;  alpha(t,u) = logaddexp(alpha(t-1,u) + lp_blank(t-1,u), alpha(t,u-1) + lp_label(t,u-1))
        nll_b = -(alpha(T_b-1, U_b) + lp_blank(T_b-1, U_b)),   loss = mean_b nll_b.
Pinning: the joint against the REAL reference module (tests/golden/make_golden_rnnt.py imports models/joint_networks.py, fixture
tests/golden/rnnt_joint_small.pt); the loss against torchaudio.functional.rnnt_loss (an independent third-party implementation of the
same definition, installed in the authoring container; its values are stored in the same fixture).  The reference's own warp_rnnt call
cannot be run: parity with warp_rnnt itself is UNPINNED, parity with the published definition is pinned through torchaudio."""
import torch


def joint_forward(sd, f, g, prefix="joint_network."):
    """reference models/joint_networks.py:80-105 (training / eval-loss form, joint_mode 'sum', act 'tanh')."""
    w = lambda n: sd[prefix + n].to(f.dtype)
    fe = torch.nn.functional.linear(f, w("linear_encoder.weight"), w("linear_encoder.bias"))          # (B, T, J)
    gd = torch.nn.functional.linear(g, w("linear_decoder.weight"), w("linear_decoder.bias"))          # (B, U+1, J)
    joint = torch.tanh(fe.unsqueeze(2) + gd.unsqueeze(1))                                             # (B, T, U+1, J)
    return torch.nn.functional.linear(joint, w("linear_joint.weight"), w("linear_joint.bias"))       # (B, T, U+1, V)


def lattice_log_probs(logits, labels, blank=0):
    """log_softmax over the vocabulary, then the two values per lattice node the loss reads: (lp_blank, lp_label), both (B, T, U+1);
    lp_label[..., u] = log p(labels[:, u]) for u < U (the last column is unused)."""
    lp = torch.log_softmax(logits, dim=-1)
    B, T, U1, V = lp.shape
    idx = torch.cat([labels.long(), torch.zeros(B, 1, dtype=torch.long)], dim=1)                     # (B, U+1)
    lp_label = lp.gather(-1, idx[:, None, :, None].expand(B, T, U1, 1)).squeeze(-1)
    return lp[..., blank], lp_label


def rnnt_loss(logits, labels, frame_len, label_len, blank=0):
    """Per-utterance negative log likelihoods (B,) and their mean; differentiable torch ops in the dtype of `logits` (use fp64)."""
    lp_blank, lp_label = lattice_log_probs(logits, labels, blank)
    out = []
    for b in range(logits.shape[0]):
        Tb, Ub = int(frame_len[b]), int(label_len[b])
        alpha = [[None] * (Ub + 1) for _ in range(Tb)]
        alpha[0][0] = logits.new_zeros(())
        for t in range(Tb):
            for u in range(Ub + 1):
                if t == 0 and u == 0:
                    continue
                terms = []
                if t > 0:
                    terms.append(alpha[t - 1][u] + lp_blank[b, t - 1, u])
                if u > 0:
                    terms.append(alpha[t][u - 1] + lp_label[b, t, u - 1])
                alpha[t][u] = torch.logsumexp(torch.stack(terms), 0)
        out.append(-(alpha[Tb - 1][Ub] + lp_blank[b, Tb - 1, Ub]))
    per = torch.stack(out)
    return per.mean(), per
