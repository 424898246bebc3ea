"""CPU: the cached-feature input pipeline (reference models/cache_loader.py:13-144): FeaturePack
round trip, fp16 storage -> numeric_type up-cast, keypoint scaling by the view's `scales`, fixed-length padding with
the reference's padding rules, collation."""
import numpy as np
import torch

from gluefactory_b200.cache_loader import CacheLoader, FeaturePack, pad_local_features, pad_to_length


def _records(rs, names, D=32):
    recs = {}
    for i, n in enumerate(names):
        k = 20 + 7 * i
        recs[n] = {"keypoints": rs.uniform(0, 100, size=(k, 2)).astype(np.float16),
                   "descriptors": rs.standard_normal((k, D)).astype(np.float16),
                   "keypoint_scores": rs.uniform(0, 1, size=(k,)).astype(np.float16)}
    return recs


def test_feature_pack_round_trip(tmp_path):
    rs = np.random.RandomState(0)
    recs = _records(rs, ["a.jpg", "b.jpg", "c.jpg"])
    path = str(tmp_path / "feats.pack")
    FeaturePack.write(path, recs)
    pack = FeaturePack(path)
    for n, r in recs.items():
        assert n in pack and set(pack.keys(n)) == set(r)
        for k, a in r.items():
            got = pack.read(n, k)
            assert got.dtype == a.dtype and np.array_equal(got, a)


def test_cache_loader_pads_scales_and_collates(tmp_path):
    rs = np.random.RandomState(1)
    names = ["x/1.jpg", "x/2.jpg"]
    recs = _records(rs, names)
    path = str(tmp_path / "scene.pack")
    FeaturePack.write(path, recs)
    loader = CacheLoader({"path": str(tmp_path / "{scene}.pack"), "add_data_path": False, "padding_fn": "pad_local_features",
                          "padding_length": 64, "numeric_type": "float32"})
    scales = torch.tensor([[2.0, 0.5], [1.0, 1.0]])
    torch.manual_seed(0)
    out = loader({"name": names, "scene": ["scene", "scene"], "scales": scales})
    assert out["keypoints"].shape == (2, 64, 2) and out["descriptors"].shape == (2, 64, 32)
    assert out["keypoint_scores"].shape == (2, 64) and out["keypoints"].dtype == torch.float32
    for i, n in enumerate(names):
        k = recs[n]["keypoints"].shape[0]
        want = torch.from_numpy(recs[n]["keypoints"]).float() * scales[i]
        assert torch.equal(out["keypoints"][i, :k], want)                      # stored points, scaled (cache_loader.py:121-131)
        assert torch.equal(out["descriptors"][i, :k], torch.from_numpy(recs[n]["descriptors"]).float())
        pad = out["keypoints"][i, k:]                                          # random_c: inside the bounding box per coordinate
        assert (pad >= want.min(0).values).all() and (pad <= want.max(0).values).all()
        dpad = out["descriptors"][i, k:]                                       # random: between the descriptors' min and max
        d = torch.from_numpy(recs[n]["descriptors"]).float()
        assert (dpad >= d.min()).all() and (dpad <= d.max()).all() and dpad.std() > 0
        assert float(out["keypoint_scores"][i, k:].abs().sum()) == 0.0         # zeros
    single = CacheLoader({"path": path, "add_data_path": False, "collate": False, "numeric_type": None})
    one = single({"name": names[:1], "scales": torch.ones(1, 2, dtype=torch.float16)})
    assert one["keypoints"].dtype == torch.float16 and one["keypoints"].shape[0] == recs[names[0]]["keypoints"].shape[0]


def test_pad_to_length_modes():
    x = torch.arange(12.0).view(6, 2)
    assert pad_to_length(x, 6) is x
    z = pad_to_length(x, 9, -2, "zeros")
    assert z.shape == (9, 2) and float(z[6:].abs().sum()) == 0
    o = pad_to_length(x[:, 0], 8, -1, "ones")
    assert o.shape == (8,) and float(o[6:].sum()) == 2
    p = pad_local_features({"keypoints": x, "keypoint_scores": x[:, 0], "descriptors": torch.randn(6, 4)}, 10)
    assert p["keypoints"].shape == (10, 2) and p["descriptors"].shape == (10, 4) and p["keypoint_scores"].shape == (10,)
