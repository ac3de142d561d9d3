"""
Drop-in boundary: the REFERENCE's own inference driver — detikzify/infer/generate.py (DetikzifyGenerator,
DetikzifyPipeline, WideNode, rollout streaming), detikzify/mcts/*, detikzify/util/{functools,generation}.py, loaded from
/root/reference and executed unmodified — runs on top of the objects ``detikzify_b200`` returns (model with ``generate``,
processor, tokenizer), with a scripted engine standing in for the GPU. This is the claim "only the ``load`` import changes".

Stubbed because they are absent offline and outside the path: torchmetrics (base class only), the TeX toolchain
(``infer/tikz.py``: pdf2image / pdfCropMargins / pymupdf → a TikzDocument that "compiles" everything), ``util/image.py``
(pymupdf, requests → two small PIL helpers), ``model/adapter`` (``has_adapter`` → False), POT's ``emd2`` (v1 models pool
with "cos"). The reference's ``evaluate/imagesim.py`` (SelfSim reward) is loaded for real on a minimal ``torchmetrics.Metric``.
Skipped on boxes without the reference checkout (the GPU box).
"""
import importlib.util
import os
import sys
import types

import pytest
import torch
from PIL import Image, ImageDraw

from scripted_engine import ScriptedEngine

REF = "/root/reference/detikzify"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not available on this box")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def reference_infer():
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("torchmetrics", "ot", "detikzify")}
    for k in list(saved):
        del sys.modules[k]
    try:
        tm = types.ModuleType("torchmetrics")
        tm.Metric = type("Metric", (), {})
        sys.modules["torchmetrics"] = tm
        for pkg in ("detikzify", "detikzify.infer", "detikzify.mcts", "detikzify.util", "detikzify.model", "detikzify.evaluate"):
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
        # real reference modules
        _load("detikzify.mcts.node", f"{REF}/mcts/node.py")
        _load("detikzify.mcts.montecarlo", f"{REF}/mcts/montecarlo.py")
        fn = _load("detikzify.util.functools", f"{REF}/util/functools.py")
        gn = _load("detikzify.util.generation", f"{REF}/util/generation.py")
        util = sys.modules["detikzify.util"]
        for mod in (fn, gn):
            for k, v in vars(mod).items():
                if not k.startswith("_"):
                    setattr(util, k, v)
        util.load = lambda image: image.convert("RGB") if isinstance(image, Image.Image) else Image.open(image).convert("RGB")

        def expand(image, size, do_trim=False):
            canvas = Image.new("RGB", (size, size), "white")
            canvas.paste(image, ((size - image.width) // 2, (size - image.height) // 2))
            return canvas
        util.expand = expand
        # stubs for what is absent offline / outside the path
        adapter = types.ModuleType("detikzify.model.adapter")
        adapter.has_adapter = lambda model: False
        sys.modules[adapter.__name__] = adapter
        adapter.AdapterProcessor = type("AdapterProcessor", (), {})
        adapter.CrossAttentionAdapterMixin = type("CrossAttentionAdapterMixin", (), {})
        _load("detikzify.util.torch", f"{REF}/util/torch.py")
        util.infer_device = sys.modules["detikzify.util.torch"].infer_device
        # torchmetrics.Metric: the slice of its protocol the reference's ImageSim relies on (states, reset, device/dtype)
        class Metric(torch.nn.Module):
            def __init__(self, **kwargs):
                super().__init__()
                self._defaults, self._dtype = {}, torch.float32

            def add_state(self, name, default, dist_reduce_fx=None):
                self._defaults[name] = default
                setattr(self, name, default.clone())

            def reset(self):
                for k, v in self._defaults.items():
                    setattr(self, k, v.clone())

            def set_dtype(self, dtype):
                self._dtype = dtype
                return self

            device = property(lambda self: self._device)
            dtype = property(lambda self: self._dtype)
        tm.Metric = Metric
        tmf = types.ModuleType("torchmetrics.functional")
        tmf.pairwise_cosine_similarity = lambda a, b: torch.nn.functional.normalize(a, dim=-1) @ torch.nn.functional.normalize(b, dim=-1).T
        sys.modules["torchmetrics.functional"] = tmf
        ot, otlp = types.ModuleType("ot"), types.ModuleType("ot.lp")
        def emd2(M, a, b):   # POT's ot.lp.emd2 (absent offline) restated: the transport LP, empty marginals = uniform
            import numpy as np
            from scipy.optimize import linprog
            M = np.asarray(M, dtype=np.float64)
            n, m = M.shape
            a = np.full(n, 1.0 / n) if len(a) == 0 else np.asarray(a, dtype=np.float64)
            b = np.full(m, 1.0 / m) if len(b) == 0 else np.asarray(b, dtype=np.float64)
            A_eq = np.zeros((n + m, n * m))
            for i in range(n):
                A_eq[i, i * m:(i + 1) * m] = 1.0
            for j in range(m):
                A_eq[n + j, j::m] = 1.0
            res = linprog(M.reshape(-1), A_eq=A_eq, b_eq=np.concatenate([a, b]), bounds=(0, None), method="highs")
            assert res.status == 0, res.message
            return float(res.fun)
        otlp.emd2 = emd2
        sys.modules["ot"], sys.modules["ot.lp"] = ot, otlp
        _load("detikzify.evaluate.imagesim", f"{REF}/evaluate/imagesim.py")
        tikz = types.ModuleType("detikzify.infer.tikz")

        class TikzDocument:
            """Stand-in for the TeX toolchain: every program 'compiles'."""
            def __init__(self, code, timeout=None):
                self.code, self.timeout = code, timeout
            is_rasterizable = True
            compiled_with_errors = False
            errors = {}

            def rasterize(self):
                return Image.new("RGB", (32, 32), "white")
        tikz.TikzDocument = TikzDocument
        sys.modules[tikz.__name__] = tikz
        yield _load("detikzify.infer.generate", f"{REF}/infer/generate.py")
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("torchmetrics", "ot", "detikzify")]:
            del sys.modules[k]
        sys.modules.update(saved)


def _ours(eos_at=40):
    from detikzify_b200.model import build_processor, preset
    from detikzify_b200.model.modeling import DetikzifyForCausalLM
    cfg = preset("tiny")
    eng = ScriptedEngine(cfg, eos_at=eos_at)
    return DetikzifyForCausalLM(cfg, engine=eng), build_processor(cfg), eng


def _figure(size=90):
    im = Image.new("RGB", (size, size + 20), "white")
    d = ImageDraw.Draw(im)
    d.line((10, 10, size - 10, size - 5), fill="black", width=3)
    d.ellipse((20, 30, 50, 60), outline="black")
    return im


def test_reference_pipeline_sample_runs_on_our_model(reference_infer):
    model, proc, eng = _ours(eos_at=30)
    pipe = reference_infer.DetikzifyPipeline(model=model, processor=proc, metric="fast")
    assert pipe.gen_kwargs["max_length"] == proc.tokenizer.model_max_length and pipe.gen_kwargs["do_sample"] is True
    doc = pipe.sample(image=_figure())
    assert isinstance(doc.code, str) and len(doc.code) > 0
    # the reference passed its own generation kwargs straight into our generate() (infer/generate.py:218-227)
    kw = eng.last_sampling
    assert kw["bad_token"] == model.config.image_token_id and kw["begin_suppress_token"] == model.config.text_config.eos_token_id
    assert kw["do_sample"] and abs(kw["temperature"] - 0.8) < 1e-6 and abs(kw["top_p"] - 0.95) < 1e-6


def test_reference_mcts_simulate_runs_on_our_model(reference_infer):
    model, proc, eng = _ours(eos_at=36)
    pipe = reference_infer.DetikzifyPipeline(model=model, processor=proc, metric="fast")
    results = list(pipe.simulate(image=_figure(), expansions=4))
    assert len(results) == 4
    for score, doc in results:
        assert score == 1 and isinstance(doc.code, str)          # scorable - compiled_with_errors with the stub compiler
    # rollouts went through the reference's ThreadPool + TokenStreamer + stopping-criteria path and our streaming contract
    assert sum(1 for c in eng.calls if c[0] == "gen_begin") >= 4


def test_reference_generator_abort_and_tree(reference_infer):
    model, proc, eng = _ours(eos_at=60)
    gen = reference_infer.DetikzifyGenerator(model=model, processor=proc, image=_figure(), metric=None,
                                             max_length=proc.tokenizer.model_max_length, temperature=0.8, top_p=0.95, top_k=0,
                                             do_sample=True)
    out = [next(gen.simulate(expansions=1)) for _ in range(2)]
    root = gen.montecarlo.root_node
    assert root.visits >= 2 and root.children and root.children[0].is_widen_node
    assert all(score == 1 for score, _ in out)
    # newline bookkeeping of the reference works with our tokenizer (vocab / decode protocol)
    assert gen.newlineinfo and all(v.num_lines >= 1 for v in gen.newlineinfo.values())


def test_reference_selfsim_metric_runs_on_our_vision_model(reference_infer):
    """metric="model": the reference's ImageSim.from_detikzify wraps OUR model.model.vision_model / image processor and
    computes the SelfSim reward from pooler_output (evaluate/imagesim.py:60-125); MCTS then min-max-normalises it."""
    model, proc, eng = _ours(eos_at=36)
    pipe = reference_infer.DetikzifyPipeline(model=model, processor=proc, metric="model")
    assert type(pipe.metric).__name__ == "ImageSim" and pipe.metric.mode == "cos"
    pipe.metric.update(img1=_figure(), img2=_figure())
    assert pipe.metric.compute() == pytest.approx(1.0)            # identical figures -> cosine 1
    pipe.metric.reset()
    results = list(pipe.simulate(image=_figure(), expansions=3))
    assert len(results) == 3 and all(-1.0 <= score <= 1.0 + 1e-9 for score, _ in results)
    assert any(c[0] == "vit_encode" for c in eng.calls)


def test_reference_emd_selfsim_agrees_with_ours(reference_infer):
    """The v2 default reward: the reference's own ImageSim in "emd" mode (evaluate/imagesim.py:105-107,121-123; POT's emd2
    restated as the transport LP) and ours (assignment solver) on the same vision model object and image processor."""
    from PIL import ImageDraw
    from detikzify_b200.evaluate.imagesim import ImageSim as Ours
    RefImageSim = sys.modules["detikzify.evaluate.imagesim"].ImageSim
    model, proc, eng = _ours(eos_at=36)
    ref = RefImageSim.from_detikzify(model, proc, mode="emd")
    ours = Ours.from_detikzify(model, proc, mode="emd")
    other = Image.new("RGB", (80, 80), "white")
    ImageDraw.Draw(other).ellipse((10, 10, 60, 70), outline="black", width=4)
    # the reference object feeds bf16 pixels (its .to(device, dtype)), ours fp32: compare the solvers on the SAME patch tokens ...
    f1, f2 = ref.get_vision_features(_figure()), ref.get_vision_features(other)
    a = ref.get_similarity(_figure(), other)
    assert f1.ndim == 2 and a == pytest.approx(Ours._emd_similarity(f1, f2), abs=1e-9) and -1.0 < a < 1.0
    # ... and the two end-to-end paths within the bf16 rounding of the inputs
    assert a == pytest.approx(ours.get_similarity(_figure(), other), abs=5e-2)
    assert ref.get_similarity(_figure(), _figure()) == pytest.approx(1.0, abs=1e-9)


def test_image_processor_matches_reference_preprocess():
    """§8 row a1: our host-side DetikzifyImageProcessor against the reference's own class
    (detikzify/model/v1/processing_detikzify.py:162-253, loaded from /root/reference; only ``timm.data`` / ``timm.models``,
    which its ``from_pretrained`` would consult for the SigLIP data config, are stubbed). The instance is built from the
    dict ``from_pretrained`` assembles (:104-117) with timm's published config of vit_so400m_patch14_siglip_384:
    input 3x384x384, mean = std = 0.5, bicubic."""
    import numpy as np
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] == "timm"}
    try:
        timm = types.ModuleType("timm"); timm.__path__ = []
        data, models = types.ModuleType("timm.data"), types.ModuleType("timm.models")
        cfg = {"input_size": [3, 384, 384], "mean": (0.5, 0.5, 0.5), "std": (0.5, 0.5, 0.5), "crop_mode": "center"}
        models.resolve_pretrained_cfg = lambda variant: types.SimpleNamespace(to_dict=lambda: dict(cfg))
        data.resolve_data_config = lambda d: dict(d)
        sys.modules.update({"timm": timm, "timm.data": data, "timm.models": models})
        ref_mod = _load("ref_processing_detikzify", f"{REF}/model/v1/processing_detikzify.py")
        ref = ref_mod.DetikzifyImageProcessor.from_pretrained("vit_so400m_patch14_siglip_384.webli")
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] == "timm"]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.modules.pop("ref_processing_detikzify", None)
    from detikzify_b200.model.processing import DetikzifyImageProcessor
    ours = DetikzifyImageProcessor(size=384)
    assert ref.size == ours.size and list(ref.image_mean) == ours.image_mean and list(ref.image_std) == ours.image_std
    assert int(ref.resample) == ours.resample == 3 and abs(ref.rescale_factor - ours.rescale_factor) < 1e-12
    rng = np.random.default_rng(3)
    images = [_figure(90), _figure(384).resize((384, 384)), Image.fromarray(rng.integers(0, 255, (200, 311, 3), dtype=np.uint8))]
    for im in images:
        a = ref(images=im, return_tensors="pt")["pixel_values"]
        b = ours(im, return_tensors="pt")["pixel_values"]
        assert a.shape == b.shape == (1, 3, 384, 384) and b.dtype == torch.float32
        assert (a.float() - b).abs().max().item() < 1e-6
