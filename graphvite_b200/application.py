"""Application layer -- the part of graphvite.application.GraphApplication the node-embedding
path needs (reference python/graphvite/application/application.py:41-187,265-453):
load / build / train / link_prediction / save_model / load_model, driven by the same keyword
arguments as the reference's config/*.yaml sections.
"""
import logging
import os
import pickle
import re

import numpy as np

from . import graph as _graph, optimizer as _optimizer, solver as _solver
from .base import auto

logger = logging.getLogger(__name__)


def link_prediction_auc(scores, labels):
    """AUC exactly as the reference computes it (application.py:444-449): sort by descending
    score, hit = cumsum(label); AUC = sum(hit[label == 0]) / (#pos * #neg)."""
    scores = np.asarray(scores)
    labels = np.asarray(labels).astype(np.int64)
    order = np.argsort(-scores, kind="stable")
    y = labels[order]
    hit = np.cumsum(y)
    total = int((y == 0).sum()) * int((y == 1).sum())
    if total == 0:
        raise ValueError("link prediction needs both positive and negative edges")
    return float(hit[y == 0].sum()) / total


def Application(type, *args, **kwargs):
    """Application(type, *args, **kwargs): factory of application.py:1371-1392; only the node-embedding
    application ("graph") is in scope."""
    if type == "graph":
        return GraphApplication(*args, **kwargs)
    if type in ("word graph", "word_graph"):
        return WordGraphApplication(*args, **kwargs)
    if type in ("knowledge graph", "knowledge_graph"):
        return KnowledgeGraphApplication(*args, **kwargs)
    raise ValueError("Unknown application `%s` (this build ships `graph`, `word graph` and `knowledge graph`)" % type)


def _resolve_gpus(gpus, knowledge_graph=False):
    """The reference drives every listed GPU (`gpus: []` = all of them) from threads of one process
    (core/solver.h:184-213).  Here one process drives one GPU and N processes form the N-GPU solver:
      * under torchrun (WORLD_SIZE > 1) rank r takes gpus[r] (or GPU LOCAL_RANK for `gpus: []`) and the solver is
        created with rank / world_size -- the YAML's `gpus: [0, 1, 2, 3]` then means what it means in the reference;
      * in a single process the whole list is handed to the solver, whose front end starts one worker process per
        listed GPU (`gpus: []` = every visible GPU).  Never a silent truncation.
    Returns (device_ids, extra solver kwargs)."""
    gpus = list(gpus)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
        if gpus and len(gpus) != world:
            raise ValueError("`gpus` lists %d devices but %d processes were launched: start one process per listed "
                             "GPU (torchrun --nproc-per-node %d)" % (len(gpus), world, len(gpus)))
        device = gpus[local] if gpus else local
        torch.cuda.set_device(device)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        return [device], dict(rank=rank, world_size=world)
    if not gpus:
        # `gpus: []` = all GPUs (core/solver.h:186-191)
        try:
            import torch
            count = torch.cuda.device_count()
        except Exception:
            count = 1
        if count > 1:
            gpus = list(range(count))
    # several GPUs in one process: the solver front end starts one worker process per GPU (graphvite_b200/multi.py)
    return gpus, {}


def _tokenize(fmt, line):
    """ApplicationMixin.tokenize (application.py:197-202) with the delimiters / comment of set_format: strip the
    delimiters off both ends, cut at the comment prefix, split at delimiter characters.  (The reference splits at
    every single delimiter, `[%s]`, so two consecutive blanks yield an empty token there and its readers then reject
    the line; runs of delimiters are one separator here, and an empty line is [] rather than [""].)"""
    delimiters, comment = fmt["delimiters"], fmt["comment"]
    line = line.strip(delimiters)
    if comment:
        start = line.find(comment)
        if start != -1:
            line = line[:start]
    line = line.strip(delimiters)
    if not line:
        return []
    return re.split("[%s]+" % re.escape(delimiters), line)


class _Model(dict):
    """What save_model pickles: a dict whose keys are also attributes, like the reference's EasyDict
    (application.py:177-187), so that `model.solver.vertex_embeddings` works on either side's files."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _get_mapping(id2name, name2id):
    """ApplicationMixin.get_mapping (application.py:189-195): stored row of every current name, or an error"""
    mapping = np.empty(len(id2name), dtype=np.int64)
    for i, name in enumerate(id2name):
        if name not in name2id:
            raise ValueError("Can't find the embedding for `%s`" % name)
        mapping[i] = name2id[name]
    return mapping


def _hyperparameters(obj, names):
    out = {}
    for name in names:
        try:
            value = getattr(obj, name)
        except Exception:
            continue
        if isinstance(value, (bool, int, float, str)):
            out[name] = value
    return out


class GraphApplication(object):
    """GraphApplication(dim, gpus=[], cpu_per_gpu=auto, gpu_memory_limit=auto, float_type, index_type)
    (application.py:265-291)."""

    def __init__(self, dim, gpus=(), cpu_per_gpu=auto, gpu_memory_limit=auto, float_type=None, index_type=None,
                 **kwargs):
        self.dim = dim
        self.gpus = list(gpus)
        self.cpu_per_gpu = cpu_per_gpu
        self.gpu_memory_limit = gpu_memory_limit
        self.graph = _graph.Graph(index_type)
        # application.py:283-286: num_sampler_per_worker = cpu_per_gpu - 1
        num_sampler_per_worker = auto if cpu_per_gpu == auto else cpu_per_gpu - 1
        device_ids, placement = _resolve_gpus(self.gpus)
        self.solver = _solver.GraphSolver(dim, float_type, index_type, device_ids, num_sampler_per_worker,
                                          gpu_memory_limit, **dict(placement, **kwargs))

    def set_format(self, delimiters=" \t\r\n", comment="#"):
        """application.py:64-76"""
        self._format = dict(delimiters=delimiters, comment=comment)
        return self

    def load(self, **kwargs):
        fmt = getattr(self, "_format", None)
        if fmt and "file_name" in kwargs:
            kwargs = dict(fmt, **kwargs)
        self.graph.load(**kwargs)
        return self

    def evaluate(self, task, **kwargs):
        """application.py:107-129: dispatch on the task name"""
        name = task.replace(" ", "_")
        if name == "link_prediction":
            result = self.link_prediction(**kwargs)
        elif name == "node_classification":
            result = self.node_classification(**kwargs)
        else:
            raise ValueError("task `%s` is not available in this build" % task)
        for metric, value in sorted(result.items()):
            logger.warning("%s: %g" % (metric, value))
        return result

    def build(self, optimizer=auto, **kwargs):
        if isinstance(optimizer, dict):  # cmd.py:99-106: a dict becomes Optimizer(**dict)
            optimizer = _optimizer.Optimizer(**optimizer)
        self.solver.build(self.graph, optimizer, **kwargs)
        return self

    def train(self, **kwargs):
        self.solver.train(**kwargs)
        return self

    # application.py:353-453, scored with the solver's own predict kernel instead of a torch module
    def link_prediction(self, H=None, T=None, Y=None, file_name=None, filter_H=None, filter_T=None,
                        filter_file=None):
        fmt = getattr(self, "_format", dict(delimiters=" \t\r\n", comment="#"))

        def read(path, width):
            rows = []
            with open(path, "r") as fin:
                for i, line in enumerate(fin):
                    tokens = _tokenize(fmt, line)
                    if not tokens:
                        continue
                    if len(tokens) != width:
                        raise ValueError("Invalid line format at line %d in %s" % (i + 1, path))
                    rows.append(tokens)
            return rows

        if file_name:
            if not (H is None and T is None and Y is None):
                raise ValueError("Evaluation data and file should not be provided at the same time")
            rows = read(file_name, 3)
            H, T, Y = [r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows]
        if H is None or T is None or Y is None:
            raise ValueError("Either evaluation data or file should be provided")
        if filter_file:
            if not (filter_H is None and filter_T is None):
                raise ValueError("Filter data and file should not be provided at the same time")
            rows = read(filter_file, 2)
            filter_H, filter_T = [r[0] for r in rows], [r[1] for r in rows]
        elif filter_H is None:
            filter_H, filter_T = [], []
        name2id = self.graph.name2id
        filters = set()
        for h, t in zip(filter_H, filter_T):
            if h in name2id and t in name2id:
                filters.add((name2id[h], name2id[t]))
        pairs, labels = [], []
        for h, t, y in zip(H, T, Y):
            if h in name2id and t in name2id and int(y) in (0, 1):
                pair = (name2id[h], name2id[t])
                if pair not in filters:
                    pairs.append(pair)
                    labels.append(int(y))
        scores = self.solver.predict(np.asarray(pairs, dtype=np.uint32).reshape(-1, 2))
        return {"AUC": link_prediction_auc(scores, labels)}

    # application.py:293-351 + linear_classification (:456-533): one-vs-rest logistic regression on
    # frozen vertex embeddings, top-k prediction with k = number of true labels, micro / macro F1
    def node_classification(self, X=None, Y=None, file_name=None, portions=(0.02,), normalization=False, times=1,
                            patience=100):
        import torch
        if file_name:
            if not (X is None and Y is None):
                raise ValueError("Evaluation data and file should not be provided at the same time")
            X, Y = [], []
            fmt = getattr(self, "_format", dict(delimiters=" \t\r\n", comment="#"))
            with open(file_name, "r") as fin:
                for i, line in enumerate(fin):
                    tokens = _tokenize(fmt, line)
                    if not tokens:
                        continue
                    if len(tokens) != 2:
                        raise ValueError("Invalid line format at line %d in %s" % (i + 1, file_name))
                    X.append(tokens[0])
                    Y.append(tokens[1])
        if X is None or Y is None:
            raise ValueError("Either evaluation data (X, Y) or a file name should be provided")
        name2id = self.graph.name2id
        classes = {c: i for i, c in enumerate(np.unique(Y))}
        rows = [(name2id[x], classes[y]) for x, y in zip(X, Y) if x in name2id]
        nodes = np.unique([r[0] for r in rows])
        position = {node: i for i, node in enumerate(nodes)}
        labels = np.zeros((len(nodes), len(classes)), dtype=np.int64)
        for node, cls in rows:
            labels[position[node], cls] = 1
        embeddings = np.array(self.solver.vertex_embeddings[nodes], dtype=np.float32)
        if normalization:
            embeddings = embeddings / np.linalg.norm(embeddings, axis=1, keepdims=True)
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        features = torch.as_tensor(embeddings, device=device)
        targets = torch.as_tensor(labels, device=device)
        metrics = {}
        for portion in portions:
            num_train = int(len(nodes) * portion)
            macro, micro = [], []
            for _ in range(times):
                order = torch.as_tensor(np.random.permutation(len(nodes)), device=device)
                train, test = order[:num_train], order[num_train:]
                # one training example per (node, label) pair, like generate_one_vs_rest
                pairs = targets[train].nonzero()
                x = features[train][pairs[:, 0]]
                y = torch.nn.functional.one_hot(pairs[:, 1], targets.shape[1]).float()
                linear = torch.nn.Linear(features.shape[1], targets.shape[1]).to(device)
                optimizer = torch.optim.SGD(linear.parameters(), lr=1, weight_decay=2e-5, momentum=0.9)
                best_loss, best_epoch = float("inf"), -1
                for epoch in range(100000):
                    optimizer.zero_grad()
                    loss = torch.nn.functional.binary_cross_entropy_with_logits(linear(x), y)
                    loss.backward()
                    optimizer.step()
                    if loss.item() < best_loss:
                        best_loss, best_epoch = loss.item(), epoch
                    if epoch == best_epoch + patience:
                        break
                with torch.no_grad():
                    logits = linear(features[test])
                    truth = targets[test]
                    count = truth.sum(dim=1, keepdim=True)
                    ranked, _ = logits.sort(dim=1, descending=True)
                    thresholds = ranked.gather(1, (count - 1).clamp(min=0))
                    predictions = (logits >= thresholds).long()
                    tp = (predictions & truth).sum(dim=0).float()
                    macro.append((2 * tp / (truth.sum(dim=0) + predictions.sum(dim=0)).float()).mean().item())
                    micro.append((2 * tp.sum() / (truth.sum() + predictions.sum()).float()).item())
            metrics["macro-F1@%g%%" % (portion * 100)] = float(np.mean(macro))
            metrics["micro-F1@%g%%" % (portion * 100)] = float(np.mean(micro))
        return metrics

    # application.py:145-187 / 131-143
    def save_model(self, file_name, save_hyperparameter=False):
        model = _Model()
        model.graph = _Model(name2id=dict(self.graph.name2id.items()), id2name=list(self.graph.id2name))
        model.solver = _Model(vertex_embeddings=np.array(self.solver.vertex_embeddings),
                              context_embeddings=np.array(self.solver.context_embeddings))
        if save_hyperparameter:  # application.py:179-184: every int / float / str attribute
            model.graph.update(_hyperparameters(self.graph, ("num_vertex", "num_edge", "as_undirected", "normalization")))
            model.solver.update(_hyperparameters(self.solver, sorted(_solver._INT_ATTRIBUTES | _solver._FLOAT_ATTRIBUTES
                                                                     | {"model", "resume", "dim"})))
            optimizer = self.solver.optimizer
            if optimizer is not None:
                model.solver.optimizer = _Model(_hyperparameters(optimizer, [n for n in dir(optimizer)
                                                                             if not n.startswith("_")]))
                model.solver.optimizer.schedule = optimizer.schedule.type
        with open(file_name, "wb") as fout:
            pickle.dump(model, fout, protocol=pickle.HIGHEST_PROTOCOL)

    def load_model(self, file_name):
        with open(file_name, "rb") as fin:
            model = pickle.load(fin)
        mapping = _get_mapping(self.graph.id2name, model["graph"]["name2id"])  # raises on a missing name
        self.solver.vertex_embeddings[:] = np.asarray(model["solver"]["vertex_embeddings"])[mapping]
        self.solver.context_embeddings[:] = np.asarray(model["solver"]["context_embeddings"])[mapping]
        return self


class WordGraphApplication(GraphApplication):
    """WordGraphApplication(dim, gpus=[], ...) (application.py:536-573): node embeddings of a word co-occurrence
    graph; `load(file_name=corpus, window=5, min_count=5, ...)` builds the graph from a corpus."""

    def __init__(self, dim, gpus=(), cpu_per_gpu=auto, gpu_memory_limit=auto, float_type=None, index_type=None,
                 **kwargs):
        GraphApplication.__init__(self, dim, gpus, cpu_per_gpu, gpu_memory_limit, float_type, index_type, **kwargs)
        self.graph = _graph.WordGraph(index_type)


class KnowledgeGraphApplication(object):
    """KnowledgeGraphApplication(dim, gpus=[], cpu_per_gpu=auto, gpu_memory_limit=auto, float_type, index_type)
    (application.py:576-1067): load / build / train / link_prediction / entity_prediction / save_model /
    load_model for TransE, DistMult, ComplEx, SimplE, RotatE and QuatE.  Triplets are scored with the solver's own
    predict kernel (the reference's "graphvite" backend, application.py:828-853)."""

    SAMPLE_PER_DIMENSION = 7  # application.py:626

    def __init__(self, dim, gpus=(), cpu_per_gpu=auto, gpu_memory_limit=auto, float_type=None, index_type=None,
                 **kwargs):
        self.dim = dim
        self.gpus = list(gpus)
        self.cpu_per_gpu = cpu_per_gpu
        self.gpu_memory_limit = gpu_memory_limit
        self.graph = _graph.KnowledgeGraph(index_type)
        num_sampler_per_worker = auto if cpu_per_gpu == auto else cpu_per_gpu - 1  # application.py:632-638
        device_ids, placement = _resolve_gpus(self.gpus, knowledge_graph=True)
        self.solver = _solver.KnowledgeGraphSolver(dim, float_type, index_type, device_ids, num_sampler_per_worker,
                                                   gpu_memory_limit, **dict(placement, **kwargs))

    def set_format(self, delimiters=" \t\r\n", comment="#"):
        self._format = dict(delimiters=delimiters, comment=comment)
        return self

    def load(self, **kwargs):
        fmt = getattr(self, "_format", None)
        if fmt and "file_name" in kwargs:
            kwargs = dict(fmt, **kwargs)
        self.graph.load(**kwargs)
        return self

    def build(self, optimizer=auto, **kwargs):
        if isinstance(optimizer, dict):
            optimizer = _optimizer.Optimizer(**optimizer)
        self.solver.build(self.graph, optimizer, **kwargs)
        return self

    def train(self, **kwargs):
        self.solver.train(**kwargs)
        return self

    def evaluate(self, task, **kwargs):
        name = task.replace(" ", "_")
        if name == "link_prediction":
            result = self.link_prediction(**kwargs)
        elif name == "entity_prediction":
            return self.entity_prediction(**kwargs)
        else:
            raise ValueError("task `%s` is not available in this build" % task)
        for metric, value in sorted(result.items()):
            logger.warning("%s: %g" % (metric, value))
        return result

    def _tokenize(self, line):
        return _tokenize(getattr(self, "_format", dict(delimiters=" \t\r\n", comment="#")), line)

    def _read_triplets(self, file_name):
        H, R, T = [], [], []
        with open(file_name, "r") as fin:
            for i, line in enumerate(fin):
                tokens = self._tokenize(line)
                if not tokens:
                    continue
                if not 3 <= len(tokens) <= 4:
                    raise ValueError("Invalid line format at line %d in %s" % (i + 1, file_name))
                H.append(tokens[0])
                R.append(tokens[1])
                T.append(tokens[2])
        return H, R, T

    def _map_names(self, H, R, T):
        """name_map (application.py:204-219): keep the triplets whose three names are known"""
        entity2id, relation2id = self.graph.entity2id, self.graph.relation2id
        rows = [(entity2id[h], relation2id[r], entity2id[t]) for h, r, t in zip(H, R, T)
                if h in entity2id and r in relation2id and t in entity2id]
        array = np.asarray(rows, dtype=np.uint32).reshape(-1, 3)
        return array[:, 0], array[:, 1], array[:, 2]

    def _batch_size(self, sample_size):
        """get_batch_size (application.py:948-961) without the psutil bound: triplets per predict() call"""
        size = int(self.SAMPLE_PER_DIMENSION * self.dim * self.graph.num_vertex / max(1, sample_size))
        return max(1, min(size, max(1, (1 << 26) // max(1, sample_size))))

    def _scores(self, H, R, T, target):
        """one-vs-rest scores [len(H) * (2 if both else 1)][num_entity] (generate_one_vs_rest, :963-977)"""
        num_entity = self.graph.num_vertex
        every = np.arange(num_entity, dtype=np.uint32)
        rows = []
        for h, r, t in zip(H, R, T):
            if target in ("head", "both"):
                rows.append(np.stack([every, np.full(num_entity, t, np.uint32), np.full(num_entity, r, np.uint32)], 1))
            if target in ("tail", "both"):
                rows.append(np.stack([np.full(num_entity, h, np.uint32), every, np.full(num_entity, r, np.uint32)], 1))
        return self.solver.predict(np.concatenate(rows)).reshape(-1, num_entity)

    def link_prediction(self, H=None, R=None, T=None, filter_H=None, filter_R=None, filter_T=None, file_name=None,
                        filter_files=None, target="both", fast_mode=None, backend="graphvite"):
        """MR, MRR, HITS@1 / 3 / 10 of filtered ranking (application.py:787-946)."""
        if target not in ("head", "tail", "both"):
            raise ValueError("Unknown target `%s`" % target)
        if file_name:
            if not (H is None and R is None and T is None):
                raise ValueError("Evaluation data and file should not be provided at the same time")
            H, R, T = self._read_triplets(file_name)
        if H is None or R is None or T is None:
            raise ValueError("Either evaluation data or file should be provided")
        if filter_files:
            if not (filter_H is None and filter_R is None and filter_T is None):
                raise ValueError("Filter data and file should not be provided at the same time")
            filter_H, filter_R, filter_T = [], [], []
            for filter_file in filter_files:
                h, r, t = self._read_triplets(filter_file)
                filter_H += h
                filter_R += r
                filter_T += t
        elif filter_H is None:
            filter_H, filter_R, filter_T = [], [], []
        total = len(H)
        H, R, T = self._map_names(H, R, T)
        logger.info("effective triplets: %d / %d" % (len(H), total))
        filter_H, filter_R, filter_T = self._map_names(filter_H, filter_R, filter_T)
        exclude_H, exclude_T = {}, {}
        for h, r, t in zip(filter_H.tolist(), filter_R.tolist(), filter_T.tolist()):
            exclude_H.setdefault((t, r), set()).add(h)
            exclude_T.setdefault((h, r), set()).add(t)
        num_sample = len(H)
        if num_sample == 0:
            raise ValueError("no evaluation triplet is known to the graph")
        fast_mode = fast_mode or num_sample
        indexes = np.random.permutation(num_sample)[:fast_mode]
        H, R, T = H[indexes], R[indexes], T[indexes]
        num_entity = self.graph.num_vertex
        batch_size = self._batch_size(num_entity * (2 if target == "both" else 1))
        rankings = []
        for i in range(0, len(H), batch_size):
            h, r, t = H[i:i + batch_size], R[i:i + batch_size], T[i:i + batch_size]
            scores = self._scores(h, r, t, target)
            row = 0
            for hh, rr, tt in zip(h.tolist(), r.tolist(), t.tolist()):
                if target in ("head", "both"):
                    mask = np.ones(num_entity, dtype=bool)
                    mask[list(exclude_H.get((tt, rr), ()))] = False
                    mask[hh] = True
                    rankings.append(int(np.sum((scores[row] >= scores[row, hh]) & mask)))
                    row += 1
                if target in ("tail", "both"):
                    mask = np.ones(num_entity, dtype=bool)
                    mask[list(exclude_T.get((hh, rr), ()))] = False
                    mask[tt] = True
                    rankings.append(int(np.sum((scores[row] >= scores[row, tt]) & mask)))
                    row += 1
        rankings = np.asarray(rankings, dtype=np.float64)
        return {"MR": float(np.mean(rankings)), "MRR": float(np.mean(1 / rankings)),
                "HITS@1": float(np.mean(rankings <= 1)), "HITS@3": float(np.mean(rankings <= 3)),
                "HITS@10": float(np.mean(rankings <= 10))}

    def entity_prediction(self, H=None, R=None, T=None, file_name=None, save_file=None, target="tail", k=10,
                          backend="graphvite"):
        """Top-k entities for the missing head / tail of every triplet (application.py:646-785): a list, per
        triplet, of (entity name, score) pairs; optionally pickled / written as text to `save_file`."""
        if target not in ("head", "tail"):
            raise ValueError("Unknown target `%s`" % target)
        if file_name:
            if not (H is None and R is None and T is None):
                raise ValueError("Evaluation data and file should not be provided at the same time")
            H, R, T = [], [], []
            with open(file_name, "r") as fin:
                for i, line in enumerate(fin):
                    tokens = self._tokenize(line)
                    if not tokens:
                        continue
                    if len(tokens) == 3:
                        h, r, t = tokens
                    elif len(tokens) == 2:  # the missing side is absent from the file
                        h, r, t = (tokens[0], tokens[1], tokens[0]) if target == "tail" else (tokens[1], tokens[0], tokens[1])
                    else:
                        raise ValueError("Invalid line format at line %d in %s" % (i + 1, file_name))
                    H.append(h)
                    R.append(r)
                    T.append(t)
        if target == "tail":
            if H is None or R is None:
                raise ValueError("Either evaluation data or file should be provided")
            T = H if T is None else T
        else:
            if T is None or R is None:
                raise ValueError("Either evaluation data or file should be provided")
            H = T if H is None else H
        H, R, T = self._map_names(H, R, T)
        id2entity = self.graph.id2entity
        k = min(k, self.graph.num_vertex)
        batch_size = self._batch_size(self.graph.num_vertex)
        recalls = []
        for i in range(0, len(H), batch_size):
            scores = self._scores(H[i:i + batch_size], R[i:i + batch_size], T[i:i + batch_size], target)
            top = np.argsort(-scores, axis=1, kind="stable")[:, :k]
            for row, index in enumerate(top):
                recalls.append([(id2entity[e], float(scores[row, e])) for e in index])
        if save_file:
            if save_file.endswith(".pkl"):
                with open(save_file, "wb") as fout:
                    pickle.dump(recalls, fout, protocol=pickle.HIGHEST_PROTOCOL)
            else:
                with open(save_file, "w") as fout:
                    for recall in recalls:
                        fout.write("\t".join("%s\t%g" % pair for pair in recall) + "\n")
        return recalls

    def save_model(self, file_name, save_hyperparameter=False):
        objects = _Model()
        objects.graph = _Model(entity2id=dict(self.graph.entity2id.items()), id2entity=list(self.graph.id2entity),
                               relation2id=dict(self.graph.relation2id.items()),
                               id2relation=list(self.graph.id2relation))
        objects.solver = _Model(entity_embeddings=np.array(self.solver.entity_embeddings),
                                relation_embeddings=np.array(self.solver.relation_embeddings))
        if save_hyperparameter:
            objects.solver.update(_hyperparameters(self.solver, sorted(_solver._KG_INT_ATTRIBUTES |
                                                                       _solver._KG_FLOAT_ATTRIBUTES | {"model", "dim"})))
        with open(file_name, "wb") as fout:
            pickle.dump(objects, fout, protocol=pickle.HIGHEST_PROTOCOL)

    def load_model(self, file_name):
        """set_parameters (application.py:640-644): rows are matched by entity / relation NAME; a name of the current
        graph that the checkpoint lacks is an error, as in the reference's get_mapping"""
        with open(file_name, "rb") as fin:
            model = pickle.load(fin)
        entity_mapping = _get_mapping(self.graph.id2entity, model["graph"]["entity2id"])
        relation_mapping = _get_mapping(self.graph.id2relation, model["graph"]["relation2id"])
        self.solver.entity_embeddings[:] = np.asarray(model["solver"]["entity_embeddings"])[entity_mapping]
        self.solver.relation_embeddings[:] = np.asarray(model["solver"]["relation_embeddings"])[relation_mapping]
        return self


__all__ = ["Application", "GraphApplication", "WordGraphApplication", "KnowledgeGraphApplication",
           "link_prediction_auc"]
