"""CPU: host-side logic of the drop-in boundary — constructor/state_dict/init parity with the
reference classes (via the committed KAT fixture), flag parsing, checkpoint naming, shard maths."""
import numpy as np
import pytest
import torch

from helpers import load_golden


def test_init_consumes_rng_like_reference():
    from cgs_b200.nets import NewCritic, UnetDecoder
    d = load_golden("kat_init_c1.npz")
    torch.manual_seed(0)
    c = NewCritic(bottleneck=32, chfak=1, dropout=0.3)
    m = UnetDecoder(bottleneck=32, chfak=1)
    x = torch.rand(4, 3, 64, 64)
    ck = [k[2:] for k in d.files if k.startswith("c.")]
    mk = [k[2:] for k in d.files if k.startswith("m.")]
    assert list(c.state_dict().keys()) == ck and list(m.state_dict().keys()) == mk
    for k, v in c.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), d["c." + k])
    for k, v in m.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), d["m." + k])
    np.testing.assert_array_equal(x.numpy(), d["x_full"])
    assert [n for n, _ in m.named_parameters()][:2] == ["dec_model.0.weight", "dec_model.0.bias"]


@pytest.mark.parametrize("K", [1, 2, 5])
def test_shapes_match_survey_table(K):
    from cgs_b200.nets import NewCritic, UnetDecoder
    import cgs_b200.synth as synth
    c, m = NewCritic(chfak=K), UnetDecoder(chfak=K)
    assert {k: tuple(v.shape) for k, v in c.state_dict().items()} == synth.critic_shapes(K)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == synth.masker_shapes(K)
    if K == 1:
        assert sum(p.numel() for p in c.parameters()) == 11873 and sum(p.numel() for p in m.parameters()) == 13785
    if K == 5:
        assert sum(p.numel() for p in c.parameters()) == 289761 and sum(p.numel() for p in m.parameters()) == 305913


def test_unsupported_variants_raise():
    from cgs_b200.nets import NewCritic, UnetDecoder
    with pytest.raises(NotImplementedError):
        NewCritic(pool="stride")
    with pytest.raises(NotImplementedError):
        UnetDecoder(upsample=False)


def test_checkpoint_names_match_reference_defaults():
    from cgs_b200.train_handler import Handler, parse_args
    H = Handler(parse_args([]), device="cpu")
    # names printed from the real Handler
    assert H.save_paths["critic"] == ("default-model/saves/critic-rewidx=1-cepochs=15-datamode=trunk-"
                                      "datasize=100000-shift=12-chfak=1-dropout=0.3.pt")
    assert H.save_paths["masker"] == "default-model/saves/masker-mepochs=1-L1=0.5-inject=True.pt"
    a = parse_args(["-frozen", "-noinject", "--chfak", "5"])
    assert a.live is False and a.inject is False and a.chfak == 5


def test_uint8_to_float_matches_float64_path():
    # segment() divides in float64 then casts (main.py:1127,1134); the kernel divides in fp32
    v = np.arange(256)
    assert np.array_equal((v / 255.0).astype(np.float32), v.astype(np.float32) / np.float32(255.0))


def test_shard_and_shift_draws():
    from cgs_b200.train_handler import Handler, parse_args
    H = Handler(parse_args([]), device="cpu", rank=1, world_size=4)
    assert H._shard(64) == slice(16, 32) and H._shard(10) == slice(3, 6)
    torch.manual_seed(3)
    r = [H._shift_roll() for _ in range(50)]
    assert all(-12 < v < 12 for v in r) and any(v < 0 for v in r) and any(v > 0 for v in r)
    torch.manual_seed(3)
    xshift = int(12 * torch.rand(1)); left = bool(torch.rand(1) > 0.5)
    assert r[0] == (xshift if left else -xshift)


def test_synthetic_generator_is_deterministic():
    import cgs_b200.synth as synth
    X1, Y1, _ = synth.synthetic_frames(64, seed=0)
    X2, Y2, _ = synth.synthetic_frames(64, seed=0)
    assert np.array_equal(X1, X2) and np.array_equal(Y1, Y2)
    assert X1.shape == (64, 64, 64, 3) and X1.dtype == np.uint8 and Y1.shape == (7, 64)
    Ys = synth.sparse_event_labels(1000, seed=1)
    assert Ys.max() <= 1.0 and Ys[1].sum() > Ys[0].sum()
