"""Drives `GraphSolver(dim, device_ids=[0, 0])` and `KnowledgeGraphSolver(dim, device_ids=[0, 0])` -- the in-process multi-GPU front end (graphvite_b200/multi.py) -- on the
CUDA emulation: two worker processes over gloo, emulated device memory.  Launched by tests/test_emulated_multi_frontend.py
with GV_EMULATE=1 and PYTHONPATH pointing at tests/emu/_pkg (the spawned workers inherit both)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOY = os.path.join(ROOT, "tests", "golden", "toy_graph.txt")


def main():
    import graphvite_b200 as gv
    from graphvite_b200.multi import SpawnedGraphSolver
    graph = gv.graph.Graph()
    graph.load(TOY)
    solver = gv.solver.GraphSolver(32, device_ids=[0, 0], num_sampler_per_worker=1)
    assert isinstance(solver, SpawnedGraphSolver)
    solver.build(graph, gv.optimizer.SGD(0.025, 0.005), num_negative=1, batch_size=300, episode_size=2)
    assert solver.num_partition == 2 and solver.num_worker == 2 and solver.batch_size == 300
    solver.train("LINE", num_epoch=6, augmentation_step=2, random_walk_length=6, random_walk_batch_size=10)
    vertex, context = solver.vertex_embeddings, solver.context_embeddings
    assert vertex.shape == (graph.num_vertex, 32) and np.isfinite(vertex).all() and np.abs(context).sum() > 0
    first = float(np.linalg.norm(vertex))
    pairs = np.array([[0, 1], [2, 3], [5, 7]], dtype=np.uint32)
    np.testing.assert_allclose(solver.predict(pairs), np.einsum("ij,ij->i", vertex[pairs[:, 0]], context[pairs[:, 1]]),
                               rtol=1e-4, atol=1e-6)
    solver.vertex_embeddings[0] *= 2  # an edit through the view reaches the workers
    np.testing.assert_allclose(solver.predict(pairs)[0], float(np.dot(vertex[0], context[1])), rtol=1e-4, atol=1e-6)
    solver.train("LINE", num_epoch=3, resume=True, augmentation_step=2, random_walk_length=6, random_walk_batch_size=10)
    assert solver.resume and float(np.linalg.norm(solver.vertex_embeddings)) != first
    # an edge-list graph and an array graph travel as recipes too
    other = gv.graph.Graph()
    other.load([(str(i), str((i * 7 + 1) % 40)) for i in range(40)] * 3)
    solver.build(other, gv.optimizer.Adam(1e-3), num_negative=2, batch_size=60, episode_size=2)
    solver.train("DeepWalk", num_epoch=4, augmentation_step=2, random_walk_length=5, random_walk_batch_size=8)
    assert solver.vertex_embeddings.shape == (other.num_vertex, 32) and np.isfinite(solver.vertex_embeddings).all()
    try:
        solver.build(other, gv.optimizer.SGD(0.1, 0.0, lambda b, n: 1.0))
        raise AssertionError("an unpicklable schedule must be refused")
    except ValueError:
        pass
    solver.close()
    knowledge_graph_front_end()
    print("front end ok")


def knowledge_graph_front_end():
    """KnowledgeGraphSolver(dim, device_ids=[0, 0]): two workers, P = 4 entity partitions, triplet-list recipe"""
    import graphvite_b200 as gv
    from graphvite_b200.multi import SpawnedKnowledgeGraphSolver
    rng = np.random.RandomState(3)
    triplets = [("e%d" % rng.randint(60), "r%d" % rng.randint(5), "e%d" % rng.randint(60)) for _ in range(600)]
    graph = gv.graph.KnowledgeGraph()
    graph.load(triplets)
    solver = gv.solver.KnowledgeGraphSolver(32, device_ids=[0, 0], num_sampler_per_worker=1)
    assert isinstance(solver, SpawnedKnowledgeGraphSolver)
    solver.build(graph, gv.optimizer.Adam(1e-2), num_partition=4, num_negative=4, batch_size=50, episode_size=2)
    assert solver.num_partition == 4 and solver.num_worker == 2 and solver.num_negative == 4
    solver.train("RotatE", num_epoch=5, margin=6, sample_batch_size=20, log_frequency=10)
    entity, relation = solver.entity_embeddings, solver.relation_embeddings
    assert entity.shape == (graph.num_vertex, 32) and relation.shape == (graph.num_relation, 32)
    assert np.isfinite(entity).all() and np.isfinite(relation).all() and np.abs(entity).sum() > 0
    samples = np.array([[0, 1, 0], [2, 3, 1], [5, 7, 4]], dtype=np.uint32)  # (h, t, r)
    logits = solver.predict(samples)
    single = gv.solver.KnowledgeGraphSolver(32, device_ids=[0], num_sampler_per_worker=1)  # same weights, one process
    single.build(graph, gv.optimizer.Adam(1e-2), num_negative=4, batch_size=50, episode_size=2)
    single.train("RotatE", num_epoch=0, margin=6, sample_batch_size=20)
    single.entity_embeddings[:] = entity
    single.relation_embeddings[:] = relation
    np.testing.assert_allclose(logits, single.predict(samples), rtol=1e-5, atol=1e-6)
    single.close()
    before = float(np.linalg.norm(entity))
    solver.train("RotatE", num_epoch=3, resume=True, margin=6, sample_batch_size=20, log_frequency=10)
    assert solver.resume and solver.model == "RotatE" and float(np.linalg.norm(solver.entity_embeddings)) != before
    solver.close()


if __name__ == "__main__":
    main()
