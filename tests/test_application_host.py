"""Host-side pieces of graphvite_b200.application that need no GPU: how `gpus` is resolved (never silently
truncated), the format-aware tokenizer of the evaluation readers, and the checkpoint helpers."""
import pickle

import numpy as np
import pytest

from graphvite_b200 import application as A


def test_gpu_lists_are_never_truncated(monkeypatch):
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    # node embeddings: the whole list reaches GraphSolver, whose front end starts one worker process per GPU
    assert A._resolve_gpus([0, 1, 2, 3]) == ([0, 1, 2, 3], {})
    assert A._resolve_gpus([2]) == ([2], {})
    # knowledge graphs: the same -- KnowledgeGraphSolver has the front end too
    assert A._resolve_gpus([0, 1], knowledge_graph=True) == ([0, 1], {})
    assert A._resolve_gpus([], knowledge_graph=True) == ([], {})  # no GPU here: `[]` stays the solver's default


def test_gpu_list_must_match_the_number_of_launched_processes(monkeypatch):
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("LOCAL_RANK", "1")
    with pytest.raises(ValueError, match="4 processes"):
        A._resolve_gpus([0, 1])


def test_tokenizer_follows_set_format():
    default = dict(delimiters=" \t\r\n", comment="#")
    assert A._tokenize(default, "a\tb 1 # note\n") == ["a", "b", "1"]
    assert A._tokenize(default, "# only a comment\n") == []
    custom = dict(delimiters=",;\n", comment="//")
    assert A._tokenize(custom, "x,y;1// c\n") == ["x", "y", "1"]
    assert A._tokenize(custom, "a b,c\n") == ["a b", "c"]  # a blank is not a delimiter in this format


def test_checkpoint_objects_are_attribute_accessible_and_picklable(tmp_path):
    model = A._Model()
    model.graph = A._Model(name2id={"a": 0, "b": 1}, id2name=["a", "b"])
    model.solver = A._Model(vertex_embeddings=np.arange(4, dtype=np.float32).reshape(2, 2))
    path = tmp_path / "model.pkl"
    with open(path, "wb") as fout:
        pickle.dump(model, fout)
    with open(path, "rb") as fin:
        loaded = pickle.load(fin)
    assert loaded.graph.name2id["b"] == 1 and loaded["solver"]["vertex_embeddings"][1, 1] == 3  # EasyDict-style + dict
    with pytest.raises(AttributeError):
        loaded.missing


def test_mapping_raises_on_a_name_the_checkpoint_lacks():
    np.testing.assert_array_equal(A._get_mapping(["b", "a"], {"a": 0, "b": 1}), [1, 0])
    with pytest.raises(ValueError, match="Can't find the embedding for `c`"):
        A._get_mapping(["a", "c"], {"a": 0, "b": 1})
