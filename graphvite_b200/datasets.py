"""Synthetic graphs shaped like the reference's benchmark datasets.

There is no network in the build / benchmark environment, so the datasets of
python/graphvite/dataset.py cannot be downloaded; these generators produce power-law
edge lists with the published vertex / edge counts (doc/source/benchmark.rst:20-24).
"""
import numpy as np

# |V|, number of edge lines (benchmark.rst:20; BlogCatalog figures are the well-known dataset sizes)
SHAPES = {
    "youtube": (1138499, 4945382),
    "blogcatalog": (10312, 333983),
    "toy": (300, 1500),
    # the Youtube shape with hub degrees capped at ~2000: node2vec's per-edge alias tables take Sigma deg^2 x 8 B
    # (209 GB on "youtube", which needs all 8 GPUs; 27 GB here, which fits one)
    "youtube_capped": (1138499, 4945382),
    # Friendster (65.6 M vertices, 1.806 G edge lines: BASELINE.json config 5) at 1/16 scale
    "friendster_lite": (4100000, 112875000),
    "friendster": (65600000, 1806000000),
}
MAX_DEGREE = {"youtube": 29000, "youtube_capped": 2000}
# graphs that are only ever handed over as integer arrays (graph.Graph.load_arrays), never written as text
BINARY_GRAPHS = {"friendster_lite", "friendster"}


def named_edges(name, seed=20260922, num_edge=None):
    """Edge arrays (u, v) of the synthetic graph `name`; num_edge overrides the edge count (evaluation samples of
    the same generative model)."""
    num_vertex, edges = SHAPES[name]
    return power_law_edges(num_vertex, edges if num_edge is None else num_edge, seed=seed,
                           max_degree=MAX_DEGREE.get(name), cover=num_edge is None)


def power_law_edges(num_vertex, num_edge, exponent=2.1, seed=20260922, max_degree=None, cover=True):
    """Chung-Lu style edge list: endpoints drawn proportionally to power-law weights.

    Returns int64 arrays (u, v) of length num_edge without self loops; every vertex id in
    [0, num_vertex) occurs at least once (so that the graph has exactly num_vertex vertices)."""
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, num_vertex + 1, dtype=np.float64)
    weights = ranks ** (-1.0 / (exponent - 1.0))
    if max_degree is not None:
        weights = np.minimum(weights, weights.sum() * max_degree / (2.0 * num_edge))
    cdf = np.cumsum(weights)
    cdf /= cdf[-1]
    u = np.searchsorted(cdf, rng.random(num_edge), side="right")
    v = np.searchsorted(cdf, rng.random(num_edge), side="right")
    u = np.minimum(u, num_vertex - 1)
    v = np.minimum(v, num_vertex - 1)
    # the first num_vertex edges make sure every vertex appears: edge i touches vertex perm[i]
    if cover:
        covered = min(num_vertex, num_edge)
        u[:covered] = rng.permutation(num_vertex)[:covered]
    loops = u == v
    v[loops] = (v[loops] + 1 + rng.integers(0, num_vertex - 1, loops.sum())) % num_vertex
    return u.astype(np.int64), v.astype(np.int64)


def write_edge_list(path, u, v, weights=None):
    """Write `u v [w]` lines; node names are the decimal ids."""
    with open(path, "w") as fout:
        if weights is None:
            np.savetxt(fout, np.stack([u, v], axis=1), fmt="%d", delimiter="\t")
        else:
            for a, b, w in zip(u, v, weights):
                fout.write("%d\t%d\t%g\n" % (a, b, w))


def synthetic_graph_file(name, path, seed=20260922):
    num_vertex, num_edge = SHAPES[name]
    u, v = power_law_edges(num_vertex, num_edge, seed=seed, max_degree=29000 if name == "youtube" else None)
    write_edge_list(path, u, v)
    return num_vertex, num_edge


def link_prediction_split(u, v, portions=(100, 1, 1), seed=1024):
    """Semantics of Dataset.link_prediction_split (python/graphvite/dataset.py:318-361): every
    edge goes to split i with probability portions[i]; each test split gets as many random
    non-edges (label 0) as it has true edges (label 1).  Returns (train_mask, [(h, t, y), ...])."""
    rng = np.random.RandomState(seed)
    cdf = np.cumsum(portions, dtype=np.float32) / np.sum(portions)
    which = np.searchsorted(cdf, rng.rand(len(u)))
    num_vertex = int(max(u.max(), v.max())) + 1
    edges = set(zip(u.tolist(), v.tolist()))
    tests = []
    for i in range(1, len(portions)):
        mask = which == i
        count = int(mask.sum())
        nh, nt = [], []
        while len(nh) < count:
            a = int(rng.rand() * num_vertex)
            b = int(rng.rand() * num_vertex)
            if a != b and (a, b) not in edges and (b, a) not in edges:
                nh.append(a)
                nt.append(b)
        h = np.concatenate([u[mask], np.asarray(nh, dtype=np.int64)])
        t = np.concatenate([v[mask], np.asarray(nt, dtype=np.int64)])
        y = np.concatenate([np.ones(count, dtype=np.int64), np.zeros(count, dtype=np.int64)])
        tests.append((h, t, y))
    return which == 0, tests


# entities, relations, training triplets (doc/source/benchmark.rst:112)
KG_SHAPES = {"fb15k-237": (14541, 237, 272115), "toy": (120, 7, 1400)}


def power_law_triplets(num_entity, num_relation, num_triplet, seed=20260922):
    """Synthetic knowledge graph shaped like the benchmark's: heads and tails drawn from power-law weights,
    relations from a skewed distribution; every entity and relation occurs at least once.  Returns int64 arrays
    (head, relation, tail)."""
    rng = np.random.default_rng(seed)
    weights = np.arange(1, num_entity + 1, dtype=np.float64) ** -0.8
    cdf = np.cumsum(weights) / weights.sum()
    head = np.minimum(np.searchsorted(cdf, rng.random(num_triplet), side="right"), num_entity - 1)
    tail = np.minimum(np.searchsorted(cdf, rng.random(num_triplet), side="right"), num_entity - 1)
    relation_weights = np.arange(1, num_relation + 1, dtype=np.float64) ** -1.0
    relation = rng.choice(num_relation, size=num_triplet, p=relation_weights / relation_weights.sum())
    cover = min(num_entity, num_triplet)
    head[:cover] = rng.permutation(num_entity)[:cover]
    relation[:min(num_relation, num_triplet)] = np.arange(min(num_relation, num_triplet))
    return head.astype(np.int64), relation.astype(np.int64), tail.astype(np.int64)


def synthetic_knowledge_graph_file(name, path, seed=20260922):
    num_entity, num_relation, num_triplet = KG_SHAPES[name]
    head, relation, tail = power_law_triplets(num_entity, num_relation, num_triplet, seed)
    with open(path, "w") as fout:
        np.savetxt(fout, np.stack([head, relation, tail], axis=1), fmt="/m/%d\t/r/%d\t/m/%d")
    return num_entity, num_relation, num_triplet
