"""Host-side mirror of the reference interface, on the GPU: hub entry points, transform registry, modules, wrappers,
the fused eval stack and the retrieval-evaluation entry points. Goldens come from the unmodified reference
(tools/gen_golden.py); tolerances: CLAHE bit-exact, descriptors 1e-5 relative (conv backbone: cuDNN fp32 vs the
reference's CPU convs, tolerance stated per test)."""
import io
import numpy as np
import pytest
import torch

from oracle import descriptors_np as D
from oracle import retrieval_np as R
from tests.util import MEAN, STD, golden, load_lut, synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vgg():
    from gandtr_b200 import hub
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)                        # same seed and constructor order as tools/gen_golden.py
    return hub.gem_vgg16_hedngan(pretrained=False)


def test_hub_model_surface(vgg):
    assert repr(vgg.transform) == ("Compose(\n    Pil2Numpy()\n    ApplyClahe(clip_limit=1.0, grid_size=8, colorspace=lab)\n"
                                   "    ToTensor()\n    Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], "
                                   "strict_shape=True)\n)")
    assert vgg.meta["architecture"] == "vgg16" and vgg.meta["out_channels"] == 512 and vgg.meta["pooling"] == "gem"
    assert "pool.p" in vgg.model.state_dict()
    assert vgg.network_params.runtime["data"]["mean_std"] == [MEAN, STD]
    assert vgg.stage == "eval" and not vgg.model.training


def test_hub_transform_and_descriptor_match_reference_golden(vgg):
    from PIL import Image
    from oracle import clahe_np
    g = golden("hub_vgg16_tiny.npz")
    x = vgg.transform(Image.fromarray(g["img"]))
    assert x.is_cuda and tuple(x.shape) == (3,) + g["img"].shape[:2]
    ref = clahe_np.transform_u8(g["img"], load_lut(), MEAN, STD)
    assert np.array_equal(x.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    with torch.no_grad():
        vec = vgg(x.unsqueeze(0))
    assert tuple(vec.shape) == (512, 1)
    v = vec.cpu().numpy()
    assert abs(np.linalg.norm(v) - 1.0) < 1e-5
    # conv stack differs (cuDNN vs the reference's CPU kernels): 16 conv layers of fp32 round-off
    np.testing.assert_allclose(v, g["desc"], rtol=0, atol=2e-4)


def test_modules_match_oracle():
    from gandtr_b200.network import GeM, L2N
    rs = np.random.RandomState(3)
    f = np.abs(rs.normal(0, 1, (3, 40, 9, 13))).astype(np.float32)
    for p in (3.0, 2.92):
        pool = GeM(p=p).cuda()
        y = pool(torch.from_numpy(f).cuda())
        assert tuple(y.shape) == (3, 40, 1, 1)
        np.testing.assert_allclose(y.cpu().numpy()[:, :, 0, 0], D.gem(f, p), rtol=2e-6)
        z = L2N()(y)
        np.testing.assert_allclose(z.cpu().numpy()[:, :, 0, 0], D.l2n(D.gem(f, p)), rtol=3e-6, atol=1e-8)
    assert repr(GeM()) == "GeM(p=3.0000, eps=1e-06)" and repr(L2N()) == "L2N(eps=1e-06)"


def test_wrappers_standalone_match_reference_goldens():
    from gandtr_b200.network import CirMultiscaleAggregation, CirtorchWhiten
    g = golden("descriptors.npz")
    for ci in range(2):
        fm = [g[k] for k in sorted(k for k in g.files if k.startswith("c%d_fmap" % ci))]
        p = float(g["c%d_p" % ci])
        c = fm[0].shape[1]
        per_scale = [torch.from_numpy(D.forward_descriptor(f, p)).cuda() for f in fm]          # N x D, oracle
        wh = CirtorchWhiten({"P": g["c%d_P" % ci], "m": g["c%d_m" % ci]}, int(g["c%d_dim" % ci]), device="cuda")
        for i in range(fm[0].shape[0]):
            cols = [ps[i].reshape(c, 1).clone() for ps in per_scale]                            # D x 1 per scale
            agg = CirMultiscaleAggregation.aggregate_tensor(cols, len(cols), c, p)
            assert tuple(agg.shape) == (c,)
            np.testing.assert_allclose(agg.cpu().numpy(), g["c%d_agg" % ci][i], rtol=2e-5, atol=2e-7)
            w = wh.postprocess(agg.clone(), None, None)
            assert tuple(w.shape) == (int(g["c%d_dim" % ci]),)
            np.testing.assert_allclose(w.cpu().numpy(), g["c%d_whiten" % ci][i], rtol=1e-4, atol=2e-6)
        # device-resident exponent (no host sync)
        agg2 = CirMultiscaleAggregation.aggregate_tensor([ps[0].reshape(c, 1) for ps in per_scale], len(fm), c,
                                                         torch.tensor([p], device="cuda"))
        np.testing.assert_allclose(agg2.cpu().numpy(), g["c%d_agg" % ci][0], rtol=2e-5, atol=2e-7)


def test_fused_eval_stack_equals_wrapper_by_wrapper_path(vgg):
    """{0_cirwhiten, 1_cirmultiscale} (mdir/hub/embedding.yml:24-25): the fused K2 path against (a) the same wrappers run
    one by one through the Compose protocol and (b) a float64 restatement on the same feature maps."""
    from gandtr_b200 import network as N
    rs = np.random.RandomState(9)
    c = 512
    whit = {"P": rs.normal(0, 1, (c, c)) / np.sqrt(c), "m": 0.05 * rs.rand(c, 1)}
    runtime = {"data": vgg.network_params.runtime["data"],
               "wrappers": {"train": None, "eval": {"0_cirwhiten": {"whitening": whit, "dimensions": 128},
                                                    "1_cirmultiscale": {"scales": True}}}}
    net = N.SingleNetwork(vgg.model, N.SingleNetwork.NetworkParams(vgg.network_params.model, runtime), "cuda", frozen=True)
    img = synth_image(77, 160, 224, "smooth")
    x = vgg.transform(img).unsqueeze(0)
    with torch.no_grad():
        fused = net(x)
        assert tuple(fused.shape) == (128,)
        generic = net.wrappers["eval"](x, net.forward_batch, outputmodel=net.model)
        fmaps = [net.model.feature_map(N.CirMultiscaleAggregation.interpolate(x, s)).cpu().numpy()
                 for s in net.wrappers["eval"].wrappers[1].scales]
    ref = D.descriptor_pipeline(fmaps, p=3.0, aggregate=True, msp_is_p=True, P=whit["P"], m=whit["m"], dimensions=128)[0]
    np.testing.assert_allclose(fused.cpu().numpy(), ref, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(generic.cpu().numpy(), fused.cpu().numpy(), rtol=1e-4, atol=2e-6)
    assert abs(float(fused.norm()) - 1.0) < 1e-5
    # batches are an extension over the reference (which is batch-size-1 here): D x N
    with torch.no_grad():
        both = net(torch.cat([x, x.flip(-1)]))
    assert tuple(both.shape) == (128, 2)
    np.testing.assert_allclose(both[:, 0].cpu().numpy(), fused.cpu().numpy(), rtol=2e-5, atol=5e-6)   # cuDNN algorithm depends on batch size


def test_clahepost_wrapper_matches_reference_golden():
    from gandtr_b200.network import ClahePost
    g = golden("clahe_post.npz")
    post = ClahePost("[[0.5,0.5,0.5],[0.5,0.5,0.5]]", "1.0", device="cuda")
    y = post.postprocess(torch.from_numpy(g["x0"]).cuda(), None, None)
    assert np.array_equal(y.cpu().numpy().view(np.uint32), g["y0"].view(np.uint32))


def test_retrieval_evaluation_entry_points_match_reference_golden():
    from gandtr_b200.retrieval import ShardedIndex, compute_map_and_print
    g = golden("map_eval.npz")
    q, db = torch.from_numpy(g["q"]).cuda(), torch.from_numpy(g["db"]).cuda()
    index = ShardedIndex(db)
    s, i = index.search(q, 100)
    os_, oi = R.topk(R.scores_exact(g["q"], g["db"]), 100)
    assert np.array_equal(i.cpu().numpy(), oi)
    gnd = [{"bbx": None, "easy": e[e >= 0], "hard": h[h >= 0], "junk": j[j >= 0]}
           for e, h, j in zip(g["easy"], g["hard"], g["junk"])]
    lines = []
    avg, per = compute_map_and_print("roxford5k", index, q, gnd, printer=lines.append)
    for name in ("easy", "medium", "hard"):
        assert abs(avg["map_" + name] - float(g["map_" + name])) < 1e-6            # "identical mAP to 0.1" with margin
        np.testing.assert_allclose(per["ap_" + name], g["ap_" + name], rtol=0, atol=1e-6, equal_nan=True)
    assert lines[0].startswith(">> roxford5k: mAP E: ") and "mP@k[1, 5, 10]" in lines[1]
    gnd_old = [{"ok": np.concatenate([x["easy"], x["hard"]]), "junk": x["junk"]} for x in gnd]
    gnd_old[3]["ok"] = np.array([], dtype=np.int64)
    avg_old, per_old = compute_map_and_print("tokyo", index, q, gnd_old, printer=lines.append)
    assert abs(avg_old["map"] - float(g["old_map"])) < 1e-6 and np.isnan(per_old["ap"][3])


def test_generators_build_in_stock_pytorch():
    from gandtr_b200 import hub
    gen = hub.cyclegan(pretrained=False)
    x = gen.transform(synth_image(5, 64, 64, "smooth")).unsqueeze(0)
    with torch.no_grad():
        y = gen(x)
    assert tuple(y.shape) == (1, 3, 64, 64) and float(y.abs().max()) <= 1.0
    with pytest.raises(NotImplementedError):
        hub.gem_vgg16_cyclegan(pretrained=True, weights_dir=None)


def test_inline_gan_clahe_embed_chain_stays_on_device(vgg):
    """BASELINE config 5 in miniature: generator (stock PyTorch) -> ClahePost (K1', float input) -> VGG16 -> K2.
    The CLAHE stage is checked bit-exactly against the oracle on the generator's own output."""
    from gandtr_b200 import hub
    from gandtr_b200.network import ClahePost
    from oracle import clahe_np
    torch.manual_seed(1)
    gen = hub.cyclegan(pretrained=False)
    x = gen.transform(synth_image(9, 64, 96, "smooth")).unsqueeze(0)
    meanstd = [[0.5, 0.5, 0.5], [0.5, 0.5, 0.5]]
    post = ClahePost(meanstd, 1.0, device="cuda")
    with torch.no_grad():
        fake = gen(x)                                            # [-1, 1] tanh output, normalised with mean = std = 0.5
        eq = post.postprocess(fake, None, None)
        assert eq.is_cuda and tuple(eq.shape) == tuple(fake.shape)
        ref = clahe_np.clahe_post_f32(fake[0].cpu().numpy(), load_lut(), meanstd, clip_limit=1.0)
        assert np.array_equal(eq[0].cpu().numpy().view(np.uint32), ref.view(np.uint32))
        # re-normalise for the embedding network (ImageNet statistics) and extract
        m0 = torch.tensor(MEAN, device="cuda").view(1, 3, 1, 1)
        s0 = torch.tensor(STD, device="cuda").view(1, 3, 1, 1)
        vec = vgg((eq * 0.5 + 0.5 - m0) / s0)
    assert tuple(vec.shape) == (512, 1) and abs(float(vec.norm()) - 1.0) < 1e-5


def test_meanstd_post_matches_reference_golden_and_torch_expression():
    """MeanStdPost (wrapper.py:149-179) as ONE kernel with the reference's four roundings: bit-exact against the
    unmodified reference's CPU output and against the same torch expression evaluated on the device."""
    from gandtr_b200.network import MeanStdPost, MeanStdPre
    g = golden("augment_chain.npz")
    post = MeanStdPost("[[0.5,0.4,0.3],[0.5,0.25,0.2]]", "[[0.485,0.456,0.406],[0.229,0.224,0.225]]", device="cuda")
    x = torch.from_numpy(g["ms_x"]).cuda()
    y = post.postprocess(x, None, None)
    assert np.array_equal(y.cpu().numpy().view(np.uint32), g["ms_y"].view(np.uint32))
    expr = x.mul(post.input_meanstd[1]).add(post.input_meanstd[0]).sub(post.output_meanstd[0]).div(post.output_meanstd[1])
    assert torch.equal(y, expr)
    # 3-d input, odd plane size (scalar path), list input, and the Pre variant
    x3 = torch.randn(3, 7, 9, device="cuda")
    y3 = post.postprocess([x3, x3 * 2], None, None)
    e3 = x3.mul(post.input_meanstd[1]).add(post.input_meanstd[0]).sub(post.output_meanstd[0]).div(post.output_meanstd[1])
    assert isinstance(y3, list) and torch.equal(y3[0], e3)
    pre = MeanStdPre("[[0.5,0.4,0.3],[0.5,0.25,0.2]]", "[[0.485,0.456,0.406],[0.229,0.224,0.225]]", device="cuda")
    t, meta = pre.preprocess(x3, None)
    assert meta is None and torch.equal(t, e3) and pre.postprocess(x3, None, None) is x3
    with pytest.raises(ValueError):
        MeanStdPost("[[0,0,0],[1,0,1]]", "[[0,0,0],[1,1,1]]", device="cuda")


def test_augment_wrapper_stack_matches_reference_golden():
    """The `augment` network's wrapper string of BASELINE config 5 (finetune.yml:13) built by initialize_wrappers and run
    through Compose on MetadataTensor inputs: routing (cir_ratio_pass_through), ClahePost (K1', float input) and
    MeanStdPost, bit-exact against the unmodified reference (tools/gen_golden_chain.py; the generator between the
    wrappers is the fixture's exactly reproducible stand-in)."""
    from gandtr_b200 import network as N
    g = golden("augment_chain.npz")
    compose = N.initialize_wrappers(str(g["wrappers"]), "cuda")
    assert [type(w).__name__ for w in compose.wrappers] == ["MeanStdPost", "ClahePost", "CirRatioPassThrough"]
    inputs = [N.MetadataTensor(torch.from_numpy(x.copy()), {"image_label": [str(lab)], "name": [str(nm)]})
              for x, lab, nm in zip(g["x"], g["labels"], g["names"])]
    seen = []

    def stand_in_generator(images):
        seen.append([x is not None for x in images])
        assert all(x is None or x.is_cuda for x in images)            # Compose moved the survivors to the device
        return [None if x is None else N.as_tensor(x).flip(-1) * 0.5 for x in images]
    with torch.no_grad():
        outs = compose(inputs, stand_in_generator)
    assert seen[0] == g["passed"].tolist()
    for o, ref in zip(outs, g["y"]):
        assert o.is_cuda and np.array_equal(o.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_sequential_network_chain_config5(vgg):
    """`CirSequentialNetwork(sequence='augment,embed')` (network.py:635-677,750-756; finetune.yml:5-32): generator
    SingleNetwork with the augment wrappers + descriptor network with `cirfaketuplebatch`, one object, everything between
    the two models on the device. Checked against the same chain assembled by hand from its parts."""
    from gandtr_b200 import network as N
    wr = ("meanstd_post:[[0.5,0.5,0.5],[0.5,0.5,0.5]]:[[0.485,0.456,0.406],[0.229,0.224,0.225]],"
          "clahepost:[[0.5,0.5,0.5],[0.5,0.5,0.5]]:1.0,cir_ratio_pass_through:1.0:anc")
    params = {"type": "CirSequentialNetwork", "sequence": "augment,embed",
              "augment": {"type": "SingleNetwork",
                          "model": {"architecture": "official_resnet_generator", "no_antialias": True, "no_antialias_up": True,
                                    "input_nc": 3, "output_nc": 3, "n_blocks": 2, "norm_layer": "instance"},
                          "initialize": {"weights": "normal_p2p", "seed": 0},
                          "runtime": {"frozen": True, "wrappers": wr,
                                      "data": {"transforms": "pil2np | totensor | normalize",
                                               "mean_std": [[0.5, 0.5, 0.5], [0.5, 0.5, 0.5]]}}},
              "embed": vgg}
    net = N.initialize_network(params, "cuda").eval()
    assert isinstance(net, N.CirSequentialNetwork) and net.meta == {"in_channels": 3, "out_channels": 512}
    assert [type(w).__name__ for w in net.wrappers["eval"].wrappers] == ["CirFakeTupleBatch"]   # the embed net's, re-homed
    gen = net.networks["augment"]
    tf = N.initialize_transforms("pil2np | totensor | normalize", [[0.5] * 3, [0.5] * 3], device="cuda")
    imgs = [synth_image(40 + i, 64, 96, "smooth") for i in range(3)]
    labels = ["anc", "pos", "neg"]
    # the training loop hands 4-d tensors to the network (CirFakeTupleBatch.unsqueeze, wrapper.py:286-294)
    tuple_ = N.CirFakeTupleBatch.unsqueeze([tf(im) for im in imgs])
    tuple_ = [N.MetadataTensor(t, {"image_label": lab, "name": "t%d" % i}) for i, (t, lab) in enumerate(zip(tuple_, labels))]
    try:
        with torch.no_grad():
            out = net([tuple_])                                       # one training tuple -> D x 3
            assert tuple(out.shape) == (512, 3)
            # by hand: only the anchor goes through the generator (ratio 1.0, label 'anc'); ClahePost + MeanStdPost on all
            post, clahe = gen.wrappers["eval"].wrappers[0], gen.wrappers["eval"].wrappers[1]
            xs = [tf(im).unsqueeze(0) for im in imgs]
            xs[0] = gen.model(xs[0])
            ys = [post.postprocess(clahe.postprocess(x, None, None), None, None) for x in xs]
            ref = torch.stack([vgg.model(y).reshape(-1) for y in ys], dim=1)
        # images that skip the generator take exactly the same kernels: bit-identical. The anchor's generator output is
        # not bit-reproducible between two cuDNN calls (atomics in the transposed convolutions) and CLAHE quantises it
        assert torch.equal(out[:, 1:], ref[:, 1:])
        torch.testing.assert_close(out[:, 0], ref[:, 0], rtol=1e-2, atol=2e-5)
    finally:
        vgg.wrappers = net.wrappers                                   # give the shared fixture its wrappers back
