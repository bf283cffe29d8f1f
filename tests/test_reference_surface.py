"""Drop-in surface checks against the UNMODIFIED reference module, when it was built in this
container (oracle/_ref/libgraphvite.so, `make -C oracle ref`).  The pybind module imports without a
GPU: its classes, docstring signatures, optimizer objects and the host-side Graph are all usable on
CPU, so the mirror in graphvite_b200/ is compared with the real thing rather than with a reading of it.

Skipped where the reference build is absent (the GPU box gets the prebuilt file with the snapshot;
a clone without /root/reference has nothing to compare with).
"""
import filecmp
import importlib.util
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libgraphvite.so")
TOY = os.path.join(ROOT, "tests", "golden", "toy_graph.txt")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_PATH), reason="oracle/_ref/libgraphvite.so is not built")


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("libgraphvite", REF_PATH)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    module.init_logging(module.ERROR, "", False)
    return module


@pytest.fixture(scope="module")
def gv():
    import graphvite_b200
    return graphvite_b200


def public(obj):
    return {name for name in dir(obj) if not name.startswith("_")}


def doc_signature(function):
    """[(name, default-or-None)] from the first docstring line pybind11 was given (bind.h)."""
    head = function.__doc__.strip().split("\n")[0]
    inside = re.match(r"\w+\((.*)\)$", head).group(1)
    out = []
    for item in filter(None, (s.strip() for s in inside.split(","))):
        name, _, default = item.partition("=")
        out.append((name, default if default else None))
    return out


def our_signature(function):
    out = []
    for name, p in list(inspect.signature(function).parameters.items())[1:]:
        out.append((name, None if p.default is inspect.Parameter.empty else p.default))
    return out


def same_default(ref_text, ours, gv):
    if ref_text is None or ours is None:
        return ref_text is None and ours is None
    if ref_text == "auto":
        return ours == gv.auto
    if ref_text in ("True", "False"):
        return ours is (ref_text == "True")
    if ref_text[0] in "'\"":
        return ours == ref_text[1:-1]
    return float(ref_text) == float(ours)


@pytest.mark.parametrize("method", ["build", "train", "predict", "clear"])
def test_solver_method_signatures(ref, gv, method):
    theirs = doc_signature(getattr(ref.solver.GraphSolver_128_f_j, method))
    ours = our_signature(getattr(gv.solver.GraphSolver, method))
    assert [n for n, _ in theirs] == [n for n, _ in ours]
    for (name, a), (_, b) in zip(theirs, ours):
        assert same_default(a, b, gv), (method, name, a, b)


def test_solver_attributes_are_all_present(ref, gv):
    from graphvite_b200 import solver as S
    ours = public(gv.solver.GraphSolver) | set(S._INT_ATTRIBUTES) | set(S._FLOAT_ATTRIBUTES)
    missing = public(ref.solver.GraphSolver_128_f_j) - ours
    assert not missing, missing


def test_graph_surface(ref, gv):
    assert public(ref.graph.Graph_j) <= public(gv.graph.Graph)
    theirs = doc_signature(ref.graph.Graph_j.save)
    ours = our_signature(gv.graph.Graph.save)
    assert [n for n, _ in theirs] == [n for n, _ in ours]
    for (name, a), (_, b) in zip(theirs, ours):
        assert same_default(a, b, gv), (name, a, b)
    assert gv.auto == ref.auto
    for unit in ("KiB", "MiB", "GiB"):
        assert getattr(gv, unit)(3) == getattr(ref, unit)(3)


OPTIMIZERS = {
    "SGD": ("lr", "weight_decay"),
    "Momentum": ("lr", "weight_decay", "momentum"),
    "AdaGrad": ("lr", "weight_decay", "epsilon"),
    "RMSprop": ("lr", "weight_decay", "alpha", "epsilon"),
    "Adam": ("lr", "weight_decay", "beta1", "beta2", "epsilon"),
}


@pytest.mark.parametrize("name", sorted(OPTIMIZERS))
def test_optimizer_defaults_and_positional_order(ref, gv, name):
    fields = OPTIMIZERS[name]
    theirs, ours = getattr(ref.optimizer, name)(), getattr(gv.optimizer, name)()
    assert public(theirs) <= public(ours) | {"lr", "weight_decay", "schedule"}
    assert theirs.type == ours.type
    for field in fields:
        assert np.float32(getattr(theirs, field)) == np.float32(getattr(ours, field)), field
    assert theirs.schedule.type == ours.schedule.type == "linear"
    # positional construction assigns the same fields in the same order
    values = [0.5, 0.25, 0.75, 0.875, 0.125][:len(fields)]
    theirs, ours = getattr(ref.optimizer, name)(*values), getattr(gv.optimizer, name)(*values)
    for field, value in zip(fields, values):
        assert getattr(theirs, field) == getattr(ours, field) == value, field
    theirs, ours = getattr(ref.optimizer, name)(schedule="constant"), getattr(gv.optimizer, name)(schedule="constant")
    assert theirs.schedule.type == ours.schedule.type == "constant"


def test_lr_schedules_agree(ref, gv):
    for kind in ("constant", "linear"):
        a, b = ref.optimizer.LRSchedule(kind), gv.optimizer.LRSchedule(kind)
        for batch_id, num_batch in ((0, 10), (3, 10), (9, 10), (10, 10), (1978, 1978), (12, 7)):
            assert np.float32(a.schedule_function(batch_id, num_batch)) == \
                np.float32(b.schedule_function(batch_id, num_batch)), (kind, batch_id, num_batch)
    with pytest.raises(ValueError):  # the reference CHECK-fails (aborts the process) on the same input
        gv.optimizer.LRSchedule("cosine")


def write_graph(path, rng, num_vertex, num_edge, weighted, comment):
    with open(path, "w") as out:
        if comment:
            out.write("# header line\n")
        for i in range(num_edge):
            u, v = rng.integers(0, num_vertex, 2)
            line = "n%d\tn%d" % (u, v) if i % 3 else "n%d n%d" % (u, v)
            if weighted:
                line += " %g" % (0.25 + 0.25 * rng.integers(0, 12))
            if comment and i % 17 == 0:
                line += " # trailing comment"
            out.write(line + "\n")


@pytest.mark.parametrize("as_undirected", [True, False])
@pytest.mark.parametrize("normalization", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
def test_graph_loader_matches_the_reference_object(ref, gv, tmp_path, as_undirected, normalization, weighted):
    """Same edge list through both loaders: ids in order of first appearance, edge counts, and a
    byte-identical save() in every mode (instance/graph.cuh:137-277)."""
    rng = np.random.default_rng(7 + 2 * as_undirected + normalization)
    source = str(tmp_path / "edges.txt")
    write_graph(source, rng, 300, 2500, weighted, comment=True)
    theirs, ours = ref.graph.Graph_j(), gv.graph.Graph()
    theirs.load(source, as_undirected, normalization)
    ours.load(source, as_undirected, normalization)
    assert (theirs.num_vertex, theirs.num_edge) == (ours.num_vertex, ours.num_edge)
    assert (theirs.as_undirected, theirs.normalization) == (ours.as_undirected, ours.normalization)
    assert list(theirs.id2name) == list(ours.id2name)
    for name in list(theirs.id2name)[::37]:
        assert theirs.name2id[name] == ours.name2id[name]
    for mode, (w, anonymous) in enumerate(((True, False), (False, False), (True, True), (False, True))):
        a, b = str(tmp_path / ("ref%d.txt" % mode)), str(tmp_path / ("ours%d.txt" % mode))
        theirs.save(a, w, anonymous)
        ours.save(b, w, anonymous)
        assert filecmp.cmp(a, b, shallow=False), (w, anonymous)


def test_edge_list_overloads_match_the_reference_object(ref, gv, tmp_path):
    rng = np.random.default_rng(11)
    pairs = [("v%d" % u, "v%d" % v) for u, v in rng.integers(0, 50, (400, 2))]
    triples = [(u, v, float(0.5 + 0.5 * (i % 5))) for i, (u, v) in enumerate(pairs)]
    for edges in (pairs, triples):
        for as_undirected in (True, False):
            theirs, ours = ref.graph.Graph_j(), gv.graph.Graph()
            theirs.load(edges, as_undirected, False)
            ours.load(edges, as_undirected, False)
            assert (theirs.num_vertex, theirs.num_edge) == (ours.num_vertex, ours.num_edge)
            assert list(theirs.id2name) == list(ours.id2name)
            a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
            theirs.save(a, True, False)
            ours.save(b, True, False)
            assert filecmp.cmp(a, b, shallow=False)


def test_custom_delimiters_and_comment_prefix(ref, gv, tmp_path):
    source = str(tmp_path / "csv.txt")
    with open(source, "w") as out:
        out.write("% matrix market style comment\n")
        out.write("a,b,2\nb,c,1.5 % note\nc,a,0.5\na,d,1\n")
    theirs, ours = ref.graph.Graph_j(), gv.graph.Graph()
    theirs.load(source, True, False, ",\r\n", "%")
    ours.load(source, True, False, ",\r\n", "%")
    assert (theirs.num_vertex, theirs.num_edge) == (ours.num_vertex, ours.num_edge) == (4, 4)
    assert list(theirs.id2name) == list(ours.id2name)
    a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    theirs.save(a, True, False)
    ours.save(b, True, False)
    assert filecmp.cmp(a, b, shallow=False)


def test_toy_graph_golden_is_what_the_reference_loads(ref, gv):
    theirs, ours = ref.graph.Graph_j(), gv.graph.Graph()
    theirs.load(TOY, True, False)
    ours.load(TOY, True, False)
    assert (theirs.num_vertex, theirs.num_edge) == (ours.num_vertex, ours.num_edge)
    assert list(theirs.id2name) == list(ours.id2name)


# ---- KnowledgeGraph (instance/knowledge_graph.cuh:67-284) ------------------------------------------
def write_triplets(path, rng, num_entity, num_relation, num_triplet, weighted):
    with open(path, "w") as out:
        out.write("# knowledge graph\n")
        for i in range(num_triplet):
            h, t = rng.integers(0, num_entity, 2)
            r = rng.integers(0, num_relation)
            line = "/m/%d\t/rel/%d\t/m/%d" % (h, r, t) if i % 2 else "/m/%d /rel/%d /m/%d" % (h, r, t)
            if weighted:
                line += " %g" % (0.5 + 0.5 * rng.integers(0, 6))
            out.write(line + ("  # note\n" if i % 23 == 0 else "\n"))


def kg_flatten(gv, graph):
    from graphvite_b200 import _lib
    m = _lib.lib.gv_kgraph_flatten(graph._handle, None, None, None, None, None, None)
    h, t, r = (np.zeros(m, dtype=np.uint32) for _ in range(3))
    w, vw = np.zeros(m, dtype=np.float32), np.zeros(graph.num_vertex, dtype=np.float32)
    _lib.lib.gv_kgraph_flatten(graph._handle, h.ctypes.data, t.ctypes.data, r.ctypes.data, w.ctypes.data, None,
                               vw.ctypes.data)
    return h, t, r, w, vw


def test_knowledge_graph_surface(ref, gv):
    assert public(ref.graph.KnowledgeGraph_j) <= public(gv.graph.KnowledgeGraph)
    theirs = doc_signature(ref.graph.KnowledgeGraph_j.save)
    ours = our_signature(gv.graph.KnowledgeGraph.save)
    assert [n for n, _ in theirs] == [n for n, _ in ours]
    for (name, a), (_, b) in zip(theirs, ours):
        assert same_default(a, b, gv), (name, a, b)


@pytest.mark.parametrize("normalization", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
def test_knowledge_graph_loader_matches_the_reference_object(ref, gv, tmp_path, normalization, weighted):
    rng = np.random.default_rng(3 + normalization)
    source = str(tmp_path / "triplets.txt")
    write_triplets(source, rng, 120, 9, 1500, weighted)
    theirs, ours = ref.graph.KnowledgeGraph_j(), gv.graph.KnowledgeGraph()
    theirs.load(source, normalization)
    ours.load(source, normalization)
    assert (theirs.num_vertex, theirs.num_edge, theirs.num_relation) == (ours.num_vertex, ours.num_edge,
                                                                         ours.num_relation)
    assert theirs.normalization == ours.normalization
    assert list(theirs.id2entity) == list(ours.id2entity)
    assert list(theirs.id2relation) == list(ours.id2relation)
    for name in list(theirs.id2entity)[::17]:
        assert theirs.entity2id[name] == ours.entity2id[name]
    for name in theirs.id2relation:
        assert theirs.relation2id[name] == ours.relation2id[name]
    # anonymous save of the reference = head, tail, int(weight) in flatten order: pins heads, tails and order;
    # (its third column is the weight by mistake, knowledge_graph.cuh:275)
    path = str(tmp_path / "ref.txt")
    theirs.save(path, True)
    columns = np.loadtxt(path, dtype=np.int64, ndmin=2)
    h, t, r, w, _ = kg_flatten(gv, ours)
    np.testing.assert_array_equal(columns[:, 0], h)
    np.testing.assert_array_equal(columns[:, 1], t)
    np.testing.assert_array_equal(columns[:, 2], w.astype(np.int64))  # float -> unsigned long long truncation
    assert "#entity: %d, #relation: %d" % (ours.num_vertex, ours.num_relation) in repr(ours)
    assert repr(theirs).splitlines()[-2:] == repr(ours).splitlines()[-2:]


def test_knowledge_graph_save_is_byte_identical_where_the_reference_is_well_defined(ref, gv, tmp_path):
    """weights equal to the relation ids make the reference's `relation = weight` slip invisible"""
    rng = np.random.default_rng(9)
    relations = ["r%d" % i for i in range(6)]
    triplets = [("e0", "r0", "e1", 0.0)]  # r0 first so that relation ids follow the numbering
    for r in range(1, 6):
        triplets.append(("e1", relations[r], "e0", float(r)))
    for _ in range(300):
        r = int(rng.integers(0, 6))
        triplets.append(("e%d" % rng.integers(0, 40), relations[r], "e%d" % rng.integers(0, 40), float(r)))
    theirs, ours = ref.graph.KnowledgeGraph_j(), gv.graph.KnowledgeGraph()
    theirs.load(triplets, False)
    ours.load(triplets, False)
    assert list(theirs.id2relation) == list(ours.id2relation) == relations
    for anonymous in (False, True):
        a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
        theirs.save(a, anonymous)
        ours.save(b, anonymous)
        assert filecmp.cmp(a, b, shallow=False), anonymous


def test_knowledge_graph_normalization_weights(ref, gv, tmp_path):
    """normalized weights are not exposed by the reference's Python surface except through save()'s
    integer column; compare against a float32 restatement of normalize() (knowledge_graph.cuh:95-121)"""
    rng = np.random.default_rng(21)
    triplets = [("e%d" % rng.integers(0, 30), "r%d" % rng.integers(0, 4), "e%d" % rng.integers(0, 30),
                 float(0.5 + 0.25 * rng.integers(0, 8))) for _ in range(400)]
    plain, normalized = gv.graph.KnowledgeGraph(), gv.graph.KnowledgeGraph()
    plain.load(triplets, False)
    normalized.load(triplets, True)
    h, t, r, w, vw = kg_flatten(gv, plain)
    head_w, tail_w = {}, {}
    for i in range(len(h)):
        head_w[(h[i], r[i])] = np.float32(head_w.get((h[i], r[i]), np.float32(0)) + w[i])
        tail_w[(t[i], r[i])] = np.float32(tail_w.get((t[i], r[i]), np.float32(0)) + w[i])
    expected = np.array([w[i] / np.sqrt(np.float32(head_w[(h[i], r[i])] * tail_w[(t[i], r[i])]))
                         for i in range(len(h))], dtype=np.float32)
    h2, t2, r2, w2, vw2 = kg_flatten(gv, normalized)
    np.testing.assert_array_equal(h2, h)
    np.testing.assert_array_equal(r2, r)
    np.testing.assert_allclose(w2, expected, rtol=1e-6)
    sums = np.zeros(plain.num_vertex, dtype=np.float64)
    np.add.at(sums, h2, w2.astype(np.float64))
    np.testing.assert_allclose(vw2, sums, rtol=1e-5)


def test_knowledge_graph_rejects_malformed_lines(gv, tmp_path):
    from graphvite_b200 import _lib
    source = str(tmp_path / "bad.txt")
    with open(source, "w") as out:
        out.write("a r b\nc r\n")
    with pytest.raises(_lib.GVError, match="Invalid format at line 2"):
        gv.graph.KnowledgeGraph().load(source)
    with open(source, "w") as out:
        out.write("a r b 1 extra\n")
    with pytest.raises(_lib.GVError, match="Invalid format at line 1"):
        gv.graph.KnowledgeGraph().load(source)


# ---- WordGraph (bind.h:190-234, instance/word_graph.cuh) ------------------------------------------------------------
def write_corpus(path, rng, vocabulary, lines):
    words = ["w%d" % i for i in range(vocabulary)]
    weights = np.arange(1, vocabulary + 1) ** -1.0
    weights /= weights.sum()
    with open(path, "w") as out:
        out.write("# a corpus\n\n")
        for i in range(lines):
            sentence = rng.choice(words, size=int(rng.integers(1, 25)), p=weights)
            text = " ".join(sentence)
            if i % 13 == 0:
                text = text.replace(" ", "\t", 2) + "   # trailing comment with words w1 w2"
            out.write(text + "\n")


@pytest.mark.parametrize("window,min_count,normalization", [(5, 5, False), (2, 1, False), (7, 3, True)])
def test_word_graph_loader_matches_the_reference_object(ref, gv, tmp_path, window, min_count, normalization):
    """Vocabulary (ids by first appearance, min_count filter), co-occurrence counts and -- because the out-edges of
    a vertex are emitted in the iteration order of an unordered_map -- the edge ORDER: save() must be byte-identical."""
    rng = np.random.default_rng(11 + window)
    corpus = str(tmp_path / "corpus.txt")
    write_corpus(corpus, rng, 150, 400)
    theirs, ours = ref.graph.WordGraph_j(), gv.graph.WordGraph()
    theirs.load(corpus, window, min_count, normalization)
    ours.load(corpus, window, min_count, normalization)
    assert [n for n, _ in doc_signature(theirs.load)] == [n for n, _ in our_signature(gv.graph.WordGraph.load)]
    assert (theirs.num_vertex, theirs.num_edge) == (ours.num_vertex, ours.num_edge)
    assert ours.num_vertex > 20 and ours.num_edge > 100
    assert (theirs.as_undirected, theirs.normalization) == (ours.as_undirected, ours.normalization) == (True, normalization)
    assert list(theirs.id2name) == list(ours.id2name)
    for mode, (weighted, anonymous) in enumerate(((True, False), (False, True))):
        a, b = str(tmp_path / ("ref%d.txt" % mode)), str(tmp_path / ("ours%d.txt" % mode))
        theirs.save(a, weighted, anonymous)
        ours.save(b, weighted, anonymous)
        assert filecmp.cmp(a, b, shallow=False), (weighted, anonymous)
    assert repr(ours).splitlines()[0] == repr(theirs).splitlines()[0] == "WordGraph<uint32>"
