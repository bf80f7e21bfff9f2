"""Host-side runtime logic that needs no GPU: FlatNet's flat buffers, the weight-generation counter behind the
per-engine bf16 shadows, and optimizer.state_dict() / load_state_dict() through views of the flat state buffers."""
import io

import pytest
import torch

from gemmgan_b200 import models
from gemmgan_b200.runtime import FlatNet


def _net():
    torch.manual_seed(0)
    return models.VanillaGenerator(16, [], [], [8, 8, 40])


@pytest.mark.parametrize("name", ["rms_prop", "adam", "adamw"])
def test_optimizer_state_dict_round_trip(name):
    gen = _net()
    flat = FlatNet(gen, torch.device("cpu"), name)
    mk = {"rms_prop": lambda ps: torch.optim.RMSprop(ps, lr=5e-4),
          "adam": lambda ps: torch.optim.Adam(ps, lr=5e-4, betas=(0.9, 0.99)),
          "adamw": lambda ps: torch.optim.AdamW(ps, lr=5e-4, betas=(0.9, 0.99), weight_decay=0.01)}[name]
    opt = mk(gen.parameters())
    flat.attach_optimizer(opt)
    # what the optimizer kernel does: writes the flat state buffers in place
    flat.exp_avg_sq.uniform_(0.1, 1.0)
    if flat.exp_avg is not None:
        flat.exp_avg.normal_()
    flat.step_count[0] = 7
    sd = opt.state_dict()
    assert len(sd["state"]) == len(flat.slots)                  # every trained tensor has state, like torch after a step
    key = "square_avg" if name == "rms_prop" else "exp_avg_sq"
    w = gen.final_layer.weight
    idx = [i for i, p in enumerate(gen.parameters()) if p is w][0]
    assert torch.equal(sd["state"][idx][key], opt.state[w][key]) and float(sd["state"][idx]["step"]) == 7.0
    assert opt.state[w][key].data_ptr() >= flat.exp_avg_sq.data_ptr()     # a view, not a copy
    # save -> perturb -> load: the flat buffers (what the kernel reads) get the saved values back
    buf = io.BytesIO()
    torch.save(sd, buf)
    saved_sq = flat.exp_avg_sq.clone()
    saved_m = None if flat.exp_avg is None else flat.exp_avg.clone()
    flat.exp_avg_sq.zero_()
    if flat.exp_avg is not None:
        flat.exp_avg.zero_()
    flat.step_count[0] = 0
    buf.seek(0)
    opt.load_state_dict(torch.load(buf))
    used = torch.zeros(flat.n_used, dtype=torch.bool)
    for slot, p in flat.slots.items():
        used[flat.offsets[slot]:flat.offsets[slot] + p.numel()] = True
    assert torch.equal(flat.exp_avg_sq[used], saved_sq[used]) and float(flat.step_count[0]) == 7.0
    if saved_m is not None:
        assert torch.equal(flat.exp_avg[used], saved_m[used])
    assert opt.state[w][key].data_ptr() >= flat.exp_avg_sq.data_ptr() and \
        opt.state[w][key].data_ptr() < flat.exp_avg_sq.data_ptr() + 4 * flat.n_used
    assert opt.param_groups[0]["lr"] == 5e-4


def test_generation_counter_tracks_every_kind_of_write():
    gen = _net()
    flat = FlatNet(gen, torch.device("cpu"), "adam")
    g0 = flat.poll_external_writes()
    assert flat.poll_external_writes() == g0                    # nothing happened
    assert flat.bump() == g0 + 1                                # an optimizer kernel ran
    with torch.no_grad():
        gen.final_layer.bias.add_(1.0)                          # a write through PyTorch
    g2 = flat.poll_external_writes()
    assert g2 == g0 + 2 and flat.poll_external_writes() == g2   # seen once, not re-counted; never cleared for others
    sd = {k: v.clone() for k, v in gen.state_dict().items()}
    gen.load_state_dict(sd)
    assert flat.poll_external_writes() == g2 + 1
    # the parameters are still views of the flat buffer after load_state_dict (copy_ in place)
    assert gen.final_layer.weight.data_ptr() == flat.params.data_ptr() + 4 * flat.offsets[max(flat.slots) - 1]
