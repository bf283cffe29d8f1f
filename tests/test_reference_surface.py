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
