"""Retrieval-evaluation score objects and stage entry points.

`CirDatasetAp` mirrors mdir/components/optim/score/cirscore.py:16-82 (same constructor keys, call signature, logger
protocol and `decisive_criterion`); `SCORES`, `initialize_score` mirror mdir/components/optim/score/__init__.py:3-12.
`validate` / `infer` mirror the stage functions mdir/stages/validate.py:15-39 and mdir/stages/infer.py:17-66 for the
`cirdatasetap` / `embedding` configurations (the entry chain of `perform_scenario.py eval iccv23/eval/*.yml`).

Database and query descriptors are extracted straight into HBM (extract.extract_descriptors), scored and ranked by K3,
evaluated by K4; the ndb x nq score / rank matrices of the reference do not exist here.
"""
import copy
import os
import pickle
import time

import numpy as np
import torch
import torch.distributed as dist

from . import network as N
from .extract import extract_descriptors
from .retrieval import ShardedIndex, compute_map_and_print, shard_bounds
from .transforms import initialize_transforms

__all__ = ["CirDatasetAp", "SCORES", "initialize_score", "configdataset", "validate", "print_scores", "infer", "load_network",
           "EmbeddingOutput"]

DATASETS = ["oxford5k", "paris6k", "roxford5k", "rparis6k", "247tokyo1k"]


def configdataset(dataset, dir_main):
    """mdir/external/cirtorch/datasets/testdataset.py:6-38: reads gnd_<dataset>.pkl ({imlist, qimlist, gnd})."""
    dataset = dataset.lower()
    if dataset not in DATASETS:
        raise ValueError("Unknown dataset: {}!".format(dataset))
    gnd_fname = os.path.join(dir_main, dataset, "gnd_{}.pkl".format(dataset))
    with open(gnd_fname, "rb") as f:
        cfg = pickle.load(f)
    cfg.update(gnd_fname=gnd_fname, ext=".jpg", qext=".jpg", dir_data=os.path.join(dir_main, dataset),
               dir_images=os.path.join(dir_main, dataset, "jpg"), n=len(cfg["imlist"]), nq=len(cfg["qimlist"]),
               dataset=dataset)
    cfg["im_fname"] = lambda c, i: os.path.join(c["dir_images"], c["imlist"][i] + c["ext"])
    cfg["qim_fname"] = lambda c, i: os.path.join(c["dir_images"], c["qimlist"][i] + c["qext"])
    return cfg


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


class CirDatasetAp:
    """mAP of a retrieval dataset. `dataset` is an official name (ground truth under $GANDTR_DATA_ROOT/test, as
    get_data_root()/test in the reference) or an in-memory dict {name, images, qimages, bbxs, gnd} (image entries may be
    paths, PIL images or uint8 arrays) -- the form the synthetic configs of BASELINE.json use."""

    decisive_criterion = "val/learning/score_avg:map_medium"

    def __init__(self, params):
        params = dict(params)
        self.image_size = params.pop("image_size")
        self.dataset = params.pop("dataset")
        self.transforms = initialize_transforms(params.pop("transforms"), params.pop("mean_std"))
        if isinstance(self.dataset, dict) and {"images", "qimages", "gnd"} <= self.dataset.keys():
            self.images, self.qimages = list(self.dataset["images"]), list(self.dataset["qimages"])
            self.bbxs = list(self.dataset.get("bbxs") or [None] * len(self.qimages))
            self.gnd = self.dataset["gnd"]
            self.dataset = self.dataset.get("name", "custom")
        elif isinstance(self.dataset, dict):
            raise NotImplementedError("tsv dataset descriptions (cirscore.py:26-38) need the daan file readers, which are "
                                      "outside the hot path; pass {name, images, qimages, bbxs, gnd} instead")
        else:
            cfg = configdataset(self.dataset, os.path.join(os.environ.get("GANDTR_DATA_ROOT", "data"), "test"))
            self.images = [cfg["im_fname"](cfg, i) for i in range(cfg["n"])]
            self.qimages = [cfg["qim_fname"](cfg, i) for i in range(cfg["nq"])]
            self.bbxs = [tuple(cfg["gnd"][i]["bbx"]) if cfg["gnd"][i]["bbx"] else None for i in range(cfg["nq"])]
            self.gnd = cfg["gnd"]
        assert not params, params.keys()

    def __call__(self, network, device, logger):
        t0 = time.time()
        world, rank = _world()
        # extract_vectors(network, ...) with the default ms=[1] (cirscore.py:56): single- vs multi-scale and whitening are
        # decided by the network's own eval wrappers, which extract_descriptors runs
        print(">> {}: database images...".format(self.dataset))
        vecs = extract_descriptors(network, self.images, self.image_size, self.transforms,
                                   rank=rank, world_size=world)                  # local rows of the database shard
        print(">> {}: query images...".format(self.dataset))
        same = len(self.images) == len(self.qimages) and set(self.bbxs) == {None} and \
            all(a is b for a, b in zip(self.images, self.qimages))
        if same and world == 1:
            qvecs = vecs.clone()
        else:
            qvecs = extract_descriptors(network, self.qimages, self.image_size, self.transforms, bbxs=self.bbxs)
        t1 = time.time()
        print(">> {}: Evaluating...".format(self.dataset))
        lo, _ = shard_bounds(len(self.images), world, rank)
        index = ShardedIndex(vecs, n_total=len(self.images), index_base=lo)
        # the probe-score / histogram exchanges need bit-identical queries on every rank: the stock cuDNN backbone may pick
        # different algorithms per process, so rank 0's copy is the one everybody scores
        qvecs = index.broadcast_queries(qvecs.contiguous())
        averages, scores = compute_map_and_print(self.dataset, index, qvecs, self.gnd)
        t2 = time.time()
        first_score = scores[list(scores.keys())[0]]
        logger(None, len(first_score), "dataset", {"extract_descriptors": t1 - t0, "compute_score": t2 - t1, "total": t2 - t0},
               "scalar/time")
        logger(None, len(first_score), "score_avg", averages, "scalar/score")
        assert len({len(x) for x in scores.values()}) == 1
        for i, _ in enumerate(first_score):
            logger(i, len(first_score), "score", {x: scores[x][i] for x in scores}, "scalar/score")
        return averages


SCORES = {"cirdatasetap": CirDatasetAp}


def initialize_score(params):
    return SCORES[params.pop("type")](params)


def load_network(params, device):
    """mdir/learning/__init__.py:9-13 for the cirnet SingleNetwork; also accepts an already built network."""
    if isinstance(params, N.SingleNetwork):
        return params
    params = copy.deepcopy(params)
    params.pop("type", None)
    return N.attach_transform(N.SingleNetwork.initialize(params, device))


def validate(params, data=()):
    """stages/validate.py:15-39: {network, validation, data} -> ({"eval": {key: value}},) with keys such as
    'roxford5k/validation/score_avg:map_medium'."""
    device = torch.device("cuda", torch.cuda.current_device())
    np.random.seed(0)
    torch.manual_seed(0)
    assert params.keys() == {"network", "validation", "data"}, params.keys()
    network = load_network(params["network"], device).eval()
    validation = copy.deepcopy(params["validation"]) if not isinstance(params["validation"], dict) else dict(params["validation"])
    assert validation.pop("type", "MultiCriterialValidation") == "MultiCriterialValidation"
    validation.pop("decisive_criterion", None)
    net_defaults = dict(network.network_params.runtime.get("data", {}))
    net_defaults = {"transforms": net_defaults.get("transforms", net_defaults.get("augmentations")),
                    "mean_std": net_defaults["mean_std"]}
    metadata = {}
    with torch.no_grad():
        for val, scenario in validation.items():
            scenario = dict(scenario)
            assert scenario.pop("type", "SingleValidation") == "SingleValidation"
            criterion = initialize_score({**net_defaults, **dict(scenario["criterion"])})

            def logger(iteration, size, label, value, dtype, _val=val):
                if iteration is None and isinstance(value, dict):
                    for k, v in value.items():
                        metadata["%s/validation/%s:%s" % (_val, label, k)] = v
            criterion(network, device, logger)
    return ({"eval": metadata},)


def print_scores(parameters, _data=()):
    """mdir/examples/perform_scenario.py:19-41: the scenario step that prints what `validate` returned, same labels and
    rounding (`round(100 * value, 2)`)."""
    scores = {
        "roxford5k/validation/score_avg:map_medium": "roxford.5k medium",
        "rparis6k/validation/score_avg:map_medium": "rparis.6k medium",
        "247tokyo1k/validation/score_avg:map": "247tokyo.1k",
        "val/validation/roxford5k/score_avg:map_medium": "roxford.5k medium",
        "val/validation/rparis6k/score_avg:map_medium": "rparis.6k medium",
        "val/validation/val_eccv20/score_avg:map": "validation eccv20",
    }
    losses = ["val/validation/loss_avg:dist"]
    assert parameters.keys() == {"metadata"}, parameters.keys()
    for heading, section in parameters["metadata"].items():
        print("\n%s\n" % heading.capitalize())
        for key, value in section.items():
            if key in scores:
                print("    %-20s %s" % (scores[key], round(100 * value, 2)))
            for loss in losses:
                if loss in key:
                    print("    %-20s %s" % (key.split(":")[-1], round(float(value.tolist()), 8)))
        print()
    return ({},)


class EmbeddingOutput:
    """mdir/components/data/output.py:118-156: (image ids, n x D float64 NumPy array)."""

    def __init__(self, images):
        self.images, self.vecs = images, None

    def add_all(self, descs):
        self.vecs = descs.to(torch.float64).cpu().numpy()

    def postprocess(self):
        return self.images, self.vecs if self.vecs is not None else []


def infer(params, data):
    """stages/infer.py:17-66 for the 'embedding' output: data = (image list,) -> (metadata, names, n x D float64)."""
    device = torch.device("cuda", torch.cuda.current_device())
    np.random.seed(0)
    torch.manual_seed(0)
    images = data[0]
    output = EmbeddingOutput(images)
    if not len(images):
        return ({"status": "skipped"},) + output.postprocess()
    network = load_network(params["network"], device).eval()
    data_params = {**network.network_params.runtime.get("data", {}), **params.get("data", {}).get("test", {})}
    transform = initialize_transforms(data_params.get("transforms", data_params.get("augmentations")), data_params["mean_std"])
    t0 = time.time()
    with torch.no_grad():
        vecs = []
        # the network's wrappers decide single- vs multi-scale and whitening, as in `out = network(indata)` (infer.py:56)
        from .extract import load_image
        for item in images:
            x = transform(load_image(item, data_params.get("image_size"))).unsqueeze(0)
            out = network(x)
            vecs.append(out.reshape(-1))
        output.add_all(torch.stack(vecs))
    return ({"stats": {"images": len(images), "time": round(time.time() - t0, 2)}},) + output.postprocess()


def infer_and_learn_whitening(params, data=()):
    """mdir/stages/multistep.py:8-43: infer descriptors of a training set, learn whitening on them, optionally store it.
    params["whitening"] = {type: 'lw' | 'pca', dataset_pkl: path or {cids, qidxs, pidxs[, images]}, directory: path | None}."""
    from . import whiten as W
    assert not data
    params = dict(params)
    whitening = params.pop("whitening")
    assert whitening.keys() == {"type", "dataset_pkl", "directory"}
    pkl = whitening["dataset_pkl"]
    tag = "memory"
    if not isinstance(pkl, dict):
        tag = str(pkl).rsplit("/", 1)[-1].split("-", 1)[0]
        with open(pkl, "rb") as f:
            pkl = pickle.load(f)
    path = None
    if whitening["directory"]:
        path = os.path.join(whitening["directory"], "whitening", "%s-%s.pkl" % (whitening["type"], tag))
        if os.path.exists(path):
            return {"status": "skipped", "whitening_path": path}, None
        os.makedirs(os.path.dirname(path), exist_ok=True)
    paths = pkl.get("images") or ["/".join([x[-2:], x[-4:-2], x[-6:-4], x]) for x in pkl["cids"]]
    metadata_infer, _cids, descriptors = infer(params, (paths,))
    learn = {"lw": W.learn_lw_whitening, "pca": W.learn_pca_whitening}[whitening["type"]]
    qidxs, pidxs = [pkl["cids"][x] for x in pkl["qidxs"]], [pkl["cids"][x] for x in pkl["pidxs"]]
    if whitening["type"] == "lw":
        metadata_learn, whit = learn({}, (pkl["cids"], descriptors, qidxs, pidxs))
    else:
        metadata_learn, whit = learn({}, (descriptors,))
    if path:
        with open(path, "wb") as f:
            pickle.dump(whit, f)
    return {"infer": metadata_infer, "learn_whitening": metadata_learn, "whitening_path": path}, whit
