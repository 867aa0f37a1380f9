"""GPU: packed-weight checkpoints.  A model saved with save_packed and loaded into a freshly initialised
copy reproduces the original outputs bit for bit, holds no fp32 master weights, and its packed payload has the k-bit size."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _build(Q):
    F, L = Q.functions, Q.layers
    mlp = nn.Sequential(F.BinaryConnect(), L.LinearBin(256, 192), nn.BatchNorm1d(192), nn.Hardtanh(), F.BinaryConnect(),
                        L.LinearTer(192, 128), nn.Hardtanh(0., 1.), F.nnDorefaQuant(4), L.LinearDorefa(128, 96, bit_width=4),
                        F.nnQuantXnor(1), L.LinearXNOR(96, 40), L.LinearQuant(40, 10, dtype="log", fsr=1, bit_width=3))
    cnn = nn.Sequential(L.XNORConv2d(32, 32, 3, padding=1), nn.Hardtanh(0., 1.), F.nnDorefaQuant(8),
                        L.DorefaConv2d(32, 64, 3, padding=1, bit_width=8), nn.BatchNorm2d(64), nn.Hardtanh(0., 1.),
                        F.BinaryConnect(), L.BinConv2d(64, 32, 3, stride=2))
    return mlp, cnn


def test_packed_checkpoint_roundtrip(tmp_path):
    import pytorch_quantize_impls_b200 as Q
    torch.manual_seed(41)
    xs = (torch.randn(70, 256).cuda(), torch.rand(3, 32, 12, 12).cuda())
    for idx in range(2):
        net = _build(Q)[idx].cuda()
        for m in net.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
        net.eval()
        with torch.no_grad():
            ref = net(xs[idx])
        path = tmp_path / ("net%d.qtb" % idx)
        Q.save_packed(net, str(path))
        state = torch.load(str(path), map_location="cpu")
        assert state["format"] == "qtb200-packed-v1"
        # k-bit payload: e.g. LinearBin(256, 192) = 192 rows x 32 bytes
        if idx == 0:
            ent = state["layers"]["1"]
            assert ent["kind"] == "sign" and ent["packed"].dtype == torch.uint8 and ent["packed"].numel() == 192 * 32
            assert state["layers"]["8"]["packed"].numel() == 96 * 64            # DoReFa-4: 128 codes x 4 bit = 64 B per row
            assert "1.weight" not in state["state"] and "1.bias" in state["state"] and "2.running_mean" in state["state"]
        torch.manual_seed(idx + 100)
        fresh = _build(Q)[idx].cuda()                                           # different random weights
        Q.load_packed(fresh, str(path))
        fresh.eval()
        with torch.no_grad():
            y = fresh(xs[idx])
        assert torch.equal(y, ref)
        # re-export from the packed-only model is lossless
        again = Q.packed_state(fresh)
        for name, ent in state["layers"].items():
            key = "packed" if ent["packed"] is not None else "planes"          # LogLin layers keep two bf16 planes
            assert torch.equal(again["layers"][name][key], ent[key])
        with torch.no_grad(), Q.code_only_activations():
            y_fused = Q.fuse_inference(fresh)(xs[idx])          # in place: module names change ("1" -> "1.layer")
        cos = torch.nn.functional.cosine_similarity(y_fused.flatten(), ref.flatten(), dim=0).item()
        assert cos > 0.98, cos                                                  # BatchNorm folded: a few one-level code moves
        qlayers = [m for m in fresh.modules() if getattr(m, "_packed_only", None) is not None]
        assert len(qlayers) == (5 if idx == 0 else 3)
        assert all(m.weight.numel() == 0 for m in qlayers)                      # no fp32 master copy left
        with pytest.raises(RuntimeError):
            qlayers[0].train(True)


def test_load_packed_rejects_foreign_files():
    import pytorch_quantize_impls_b200 as Q
    net = _build(Q)[0].cuda()
    with pytest.raises(ValueError):
        Q.load_packed(net, {"format": "something-else"})
    st = Q.packed_state(net)
    del st["layers"]["1"]
    with pytest.raises(KeyError):
        Q.load_packed(net, st)
