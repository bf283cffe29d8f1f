"""Application layer -- the part of graphvite.application.GraphApplication the node-embedding
path needs (reference python/graphvite/application/application.py:41-187,265-453):
load / build / train / link_prediction / save_model / load_model, driven by the same keyword
arguments as the reference's config/*.yaml sections.
"""
import logging
import pickle

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
    raise ValueError("Unknown application `%s` (this build ships the node-embedding application `graph`)" % type)


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
        self.solver = _solver.GraphSolver(dim, float_type, index_type, self.gpus[:1], num_sampler_per_worker,
                                          gpu_memory_limit, **kwargs)

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
        def read(path, width):
            rows = []
            with open(path, "r") as fin:
                for line in fin:
                    tokens = line.split("#")[0].split()
                    if not tokens:
                        continue
                    if len(tokens) != width:
                        raise ValueError("Invalid line `%s`" % line.strip())
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
            with open(file_name, "r") as fin:
                for line in fin:
                    tokens = line.split("#")[0].split()
                    if tokens:
                        x, y = tokens
                        X.append(x)
                        Y.append(y)
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
        objects = {"graph": {"name2id": dict(self.graph.name2id.items()), "id2name": list(self.graph.id2name)},
                   "solver": {"vertex_embeddings": np.array(self.solver.vertex_embeddings),
                              "context_embeddings": np.array(self.solver.context_embeddings)}}
        with open(file_name, "wb") as fout:
            pickle.dump(objects, fout, protocol=pickle.HIGHEST_PROTOCOL)

    def load_model(self, file_name):
        with open(file_name, "rb") as fin:
            objects = pickle.load(fin)
        mapping = objects["graph"]["name2id"]
        name2id = self.graph.name2id
        for key in ("vertex_embeddings", "context_embeddings"):
            view = getattr(self.solver, key)
            stored = objects["solver"][key]
            for name, old in mapping.items():
                if name in name2id:
                    view[name2id[name]] = stored[old]
        return self


__all__ = ["Application", "GraphApplication", "link_prediction_auc"]
