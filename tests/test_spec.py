"""Seeded construction reproduces the reference's weights bit-for-bit (golden hashes from
tests/golden/make_golden.py) and the state_dict layout of SURVEY.md Appendix C."""
import hashlib
import json

import torch

from gan_aug_pfa_b200 import spec


def sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(str(tuple(v.shape)).encode())
        h.update(str(v.dtype).encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def test_gan_seeded_state_dicts_match_reference(golden_dir):
    gold = json.loads((golden_dir / "gan_full.json").read_text())
    torch.manual_seed(0)
    g, d = spec.default_state_dicts()
    assert list(g.keys()) == gold["keys_g"] and len(g) == 70
    assert list(d.keys()) == gold["keys_d"] and len(d) == 22
    assert sd_hash(g) == gold["sd_g_sha256"]
    assert sd_hash(d) == gold["sd_d_sha256"]
    n_g = sum(v.numel() for k, v in g.items() if "running" not in k and "num_batches" not in k)
    n_d = sum(v.numel() for k, v in d.items() if "running" not in k and "num_batches" not in k)
    assert (n_g, n_d) == (gold["n_params_g"], gold["n_params_d"]) == (41828995, 2768705)


def test_layout_details():
    g = spec.GeneratorSpec().default_state_dict()
    assert g["model.model.0.weight"].shape == (64, 3, 4, 4)
    assert g["model.model.3.weight"].shape == (128, 3, 4, 4) and g["model.model.3.bias"].shape == (3,)
    assert g["model.model.1.model.2.num_batches_tracked"].dtype == torch.int64
    d = spec.DiscriminatorSpec().default_state_dict()
    assert d["model.0.weight"].shape == (64, 6, 4, 4) and d["model.11.weight"].shape == (1, 512, 4, 4)
    assert "model.2.bias" not in d and "model.11.bias" in d


def test_siamese_seeded_state_dict_matches_reference(golden_dir):
    gold = torch.load(golden_dir / "siamese_small.pt")
    torch.manual_seed(0)
    sd = spec.SiameseSpec().default_state_dict()
    assert list(sd.keys()) == gold["keys"] and len(sd) == 194
    assert sd_hash(sd) == gold["sd_sha256"]
