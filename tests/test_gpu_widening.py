"""GPU: SURVEY section 8(f) rows -- DropoutMLP, Linear / Logistic regression with penalties, SOM + RBF network, batched
validation, stochastic_train -- against golden vectors produced by the real reference (oracle/gen_golden.py part E)."""
import os
import random

import numpy
import pytest

import mlp_oracle as mo
from conftest import GOLDEN
from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import learning_b200
    return learning_b200


@pytest.fixture(scope="module")
def wide():
    return numpy.load(os.path.join(GOLDEN, "widening.npz"))


def test_dropout_mlp_training_matches_reference(L, golden_meta, wide):
    tmap = {"softplus": L.ReluTransfer, "tanh": L.TanhTransfer, "gaussian": L.GaussianTransfer,
            "linear": L.LinearTransfer, "softmax": L.SoftmaxTransfer}
    for case in golden_meta["dropout"]:
        k, shape = case["key"], tuple(case["shape"])
        transfers = [tmap[case["hidden"]]() for _ in shape[1:-1]] + [tmap[case["out"]]()]
        random.seed(case["seed"])
        numpy.random.seed(case["seed"])
        model = L.DropoutMLP(shape, transfers=transfers,
                             error_func=L.MeanSquaredError() if case["error"] == 0 else L.CrossEntropyError(),
                             input_active_probability=0.8, hidden_active_probability=0.6)
        model.logging = False
        x, y = wide[k + "_x"], wide[k + "_y"]
        for step in range(case["steps"]):
            err = model.train_step(x, y)
            got = numpy.hstack([model._bias_vec] + [m.ravel() for m in model._weight_matrices])
            assert rel_err(err, wide[k + "_errors"][step]) <= 1e-8, (case, step)
            assert rel_err(got, wide[k + "_weights"][step]) <= 1e-8, (case, step, rel_err(got, wide[k + "_weights"][step]))
        # first activate after training rescales the weights by the hidden keep probability (mlp.py:374-384)
        assert rel_err(model.activate(x), wide[k + "_out"]) <= 1e-8, case
        assert rel_err(model.activate(x), wide[k + "_out"]) <= 1e-8   # ... and only once
    with pytest.raises(ValueError):
        from learning_b200.architecture.mlp import _get_active_neurons
        _get_active_neurons(0.0, 4)


def test_dropout_mlp_engines_agree_at_size(L, monkeypatch):
    """DropoutMLP at a size where the INT8 engine and the fused head are active (masks folded into W0 / dW0 rows and
    into the forward epilogue): identical seeds give identical masks, so three steps on the INT8 engine and on the
    plain DMMA schedule must leave the same weights."""
    rng = numpy.random.default_rng(3)
    x = rng.random((2048, 200)) * 2 - 1
    y = numpy.zeros((2048, 6))
    y[numpy.arange(2048), rng.integers(0, 6, 2048)] = 1.0
    weights = {}
    for precision in ("f64", "f64-dmma"):
        monkeypatch.setenv("OPL_PRECISION", precision)
        random.seed(11)
        numpy.random.seed(11)
        model = L.DropoutMLP((200, 96, 6), transfers=[L.ReluTransfer(), L.SoftmaxTransfer()], error_func=L.CrossEntropyError(),
                             input_active_probability=0.8, hidden_active_probability=0.6)
        model.logging = False
        errs = [model.train_step(x, y) for _ in range(3)]
        assert all(numpy.isfinite(errs))
        weights[precision] = numpy.hstack([model._bias_vec] + [m.ravel() for m in model._weight_matrices])
    assert rel_err(weights["f64"], weights["f64-dmma"]) <= 1e-10, rel_err(weights["f64"], weights["f64-dmma"])


def test_regression_models_match_reference(L, golden_meta, wide):
    for case in golden_meta["regression"]:
        k = case["key"]
        pen = None
        if case["penalty"]:
            pen = (L.L1Penalty if case["penalty"][0] == "l1" else L.L2Penalty)(penalty_weight=case["penalty"][1])
        cls = L.LinearRegressionModel if case["kind"] == "linear" else L.LogisticRegressionModel
        err_f = L.MeanSquaredError() if case["error"] == 0 else L.CrossEntropyError()
        x, y, w = wide[k + "_x"], wide[k + "_y"], wide[k + "_w"]
        random.seed(case["seed"])
        numpy.random.seed(case["seed"])
        model = cls(case["attrs"], case["outs"], error_func=err_f, penalty_func=pen)
        model.logging = False
        assert rel_err(model._weight_matrix.ravel(), wide[k + "_w0"]) == 0.0      # same seeded initial weights
        obj = model._get_obj(w.copy(), x, y)
        obj2, jac = model._get_obj_jac(w.copy(), x, y)
        assert rel_err(obj, wide[k + "_obj"]) <= 1e-10 and rel_err(obj2, wide[k + "_obj"]) <= 1e-10, case
        assert rel_err(jac, wide[k + "_jac"]) <= 1e-10, (case, rel_err(jac, wide[k + "_jac"]))
        assert rel_err(model.activate(x), wide[k + "_out"]) <= 1e-10, case
        # 10 training steps from the seeded weights (default optimizer: BFGS below 500 weights)
        model._weight_matrix = wide[k + "_w0"].copy().reshape(model._weight_matrix.shape)
        for step in range(10):
            err = model.train_step(x, y)
            assert rel_err(err, wide[k + "_errors"][step]) <= 1e-7, (case, step)
            assert rel_err(model._weight_matrix.ravel(), wide[k + "_weights"][step]) <= 1e-7, (case, step)
    with pytest.raises(ValueError):
        L.LinearRegressionModel(3, 2)._get_obj_jac(numpy.zeros(8), numpy.zeros((4, 3)), numpy.zeros((4, 5)))


def test_penalty_known_answers(L, wide):
    w = wide["pen_w"]
    assert rel_err(L.L1Penalty(0.3)(w), wide["pen_l1"][0]) <= 1e-14
    assert rel_err(L.L1Penalty(0.3).derivative(w), wide["pen_l1_d"]) <= 1e-14
    assert rel_err(L.L2Penalty(0.7)(w), wide["pen_l2"][0]) <= 1e-14
    assert rel_err(L.L2Penalty(0.7).derivative(w), wide["pen_l2_d"]) <= 1e-14
    with pytest.raises(TypeError):
        class Mine(L.error.PenaltyFunc):
            pass
        L.LinearRegressionModel(3, 2, penalty_func=Mine())


def test_batched_validation_equals_per_pattern_reference_semantics(L):
    """validation.get_error / get_accuracy (validation.py:235-254): the reference loops over patterns;
    the batched forward must give the same numbers."""
    rng = numpy.random.default_rng(3)
    shape, transfers = (6, 5, 3), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)]
    x, y = mo.random_classification(57, 6, 3, rng)
    numpy.random.seed(4)
    model = L.MLP(shape, transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError())
    w = numpy.hstack([model._bias_vec] + [m.ravel() for m in model._weight_matrices])
    out = mo.activate(w, shape, transfers, x)
    per_row_mse = numpy.mean([mo.error_value(mo.MSE, o, t) for o, t in zip(out, y)])
    per_row_ce = numpy.mean([mo.error_value(mo.CROSS_ENTROPY, o, t) for o, t in zip(out, y)])
    assert rel_err(L.validation.get_error(model, x, y), per_row_mse) <= 1e-12
    assert rel_err(L.validation.get_error(model, x, y, error_func=L.CrossEntropyError()), per_row_ce) <= 1e-12
    acc = numpy.mean(numpy.argmax(out, axis=1) == numpy.argmax(y, axis=1))
    assert L.validation.get_accuracy(model, x, y) == pytest.approx(acc, abs=0)


def test_stochastic_train_matches_reference(L, golden_traces, wide):
    """Model.stochastic_train + select_sample (base.py:39-135): 4 outer iterations of 6 BFGS steps on random iris
    subsets (random.sample order from the seeded generator), each mini-batch resident once per inner train."""
    t = golden_traces
    random.seed(7)
    numpy.random.seed(7)
    model = L.MLP((4, 6, 3), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError())
    model.logging = False
    err = model.stochastic_train(t["iris_x"], t["iris_y"], max_iterations=4, train_kwargs={'iterations': 6})
    got = numpy.hstack([model._bias_vec] + [m.ravel() for m in model._weight_matrices])
    assert rel_err(err, float(wide["st_error"])) <= 1e-6, (err, float(wide["st_error"]))
    assert rel_err(got, wide["st_weights"]) <= 1e-6, rel_err(got, wide["st_weights"])
    assert model.iteration == int(wide["st_iteration"])


def test_som_matches_reference(L, wide):
    """SOM (architecture/som.py:33-113): seeded weights, 25 passes of the competition / move loop as one resident-CTA
    kernel launch per pass, batched and single-pattern distances."""
    random.seed(90)
    numpy.random.seed(90)
    xs = wide["som_x"]
    som = L.SOM(3, 5)
    som.logging = False
    assert numpy.array_equal(som._weights, wide["som_w0"])
    assert som.train(xs, xs, iterations=25) is None          # the SOM reports no error
    assert rel_err(som._weights, wide["som_w"]) <= 1e-12, rel_err(som._weights, wide["som_w"])
    assert rel_err(som.activate(xs), wide["som_dist"]) <= 1e-12
    assert rel_err(som.activate(xs[3]), wide["som_dist_vec"]) <= 1e-12
    assert som.activate(xs[3]).shape == (5,)
    # the per-pattern path (post_pattern_callback) is the same arithmetic
    random.seed(90)
    numpy.random.seed(90)
    som2 = L.SOM(3, 5)
    som2.logging = False
    seen = []
    som2.train(xs, xs, iterations=2, post_pattern_callback=lambda model, a, b: seen.append(1))
    random.seed(90)
    numpy.random.seed(90)
    som3 = L.SOM(3, 5)
    som3.logging = False
    som3.train(xs, xs, iterations=2)
    assert len(seen) == 60 and numpy.array_equal(som2._weights, som3._weights)
    with pytest.raises(ValueError):
        som.activate(numpy.zeros((2, 2, 3)))


def test_rbf_matches_reference(L, golden_meta, wide):
    """RBF (architecture/rbf.py:36-231): SOM clustering (1000 passes), Gaussian similarity (+ scaling by the row sum),
    objective / Jacobian of the linear output layer, training traces incl. cluster_incrementally."""
    for case in golden_meta["rbf"]:
        k = case["key"]
        x, y, w = wide[k + "_x"], wide[k + "_y"], wide[k + "_w"]

        def make():
            random.seed(case["seed"])
            numpy.random.seed(case["seed"])
            m = L.RBF(case["attrs"], case["clusters"], case["outs"], variance=case["variance"],
                      scale_by_similarity=case["scale"], cluster_incrementally=case["incremental"])
            m.logging = False
            return m
        model = make()
        model._pre_train(x, y)
        assert rel_err(model._clustering_model._weights, wide[k + "_centers"]) <= 1e-11, case
        obj = model._get_obj(w.copy(), x, y)
        obj2, jac = model._get_obj_jac(w.copy(), x, y)
        assert rel_err(obj, wide[k + "_obj"]) <= 1e-10 and rel_err(obj2, wide[k + "_obj"]) <= 1e-10, case
        assert rel_err(jac, wide[k + "_jac"]) <= 1e-10, (case, rel_err(jac, wide[k + "_jac"]))
        assert rel_err(model.activate(x), wide[k + "_out"]) <= 1e-10, case
        assert rel_err(model._similarity_tensor, wide[k + "_sim"]) <= 1e-11, case
        assert rel_err(model.activate(x[2]), wide[k + "_out"][2]) <= 1e-10
        model = make()
        model._pre_train(x, y)
        for step in range(case["steps"]):
            err = model.train_step(x, y)
            got = numpy.hstack([model._bias_vec, model._weight_matrix.ravel()])
            assert rel_err(err, wide[k + "_errors"][step]) <= 1e-7, (case, step)
            assert rel_err(got, wide[k + "_weights"][step]) <= 1e-7, (case, step, rel_err(got, wide[k + "_weights"][step]))
        assert rel_err(model._clustering_model._weights, wide[k + "_centers_end"]) <= 1e-11, case
    with pytest.raises(TypeError):
        L.RBF(2, 3, 1, clustering_model=object())


def test_rbf_far_pattern_gets_uniform_similarity(L):
    """reference tests/architecture/test_rbf.py:36-60: a pattern far from every cluster has all similarities underflow;
    0/0 is replaced by the uniform vector instead of NaN."""
    random.seed(0)
    numpy.random.seed(0)
    clustering_model = L.SOM(1, 2, neighborhood=0)
    clustering_model.logging = False
    model = L.RBF(1, 2, 1, variance=1.0, scale_by_similarity=True, clustering_model=clustering_model)
    model._pre_train(numpy.array([[0.], [1.]]), numpy.array([[0.], [1.]]))
    assert numpy.allclose(model._clustering_model.activate([0]), [0, 1], atol=1e-3)
    assert not numpy.isnan(model.activate(numpy.array([1000.]))).any()
    assert numpy.allclose(model._similarity_tensor, [0.5, 0.5], atol=1e-3)
    assert not numpy.isnan(model.activate(numpy.array([[0.], [1000.]]))).any()
    assert numpy.allclose(model._similarity_tensor, [[0.73105858, 0.26894142], [0.5, 0.5]], atol=1e-3)


def test_rbf_trains(L):
    """reference tests/architecture/test_rbf.py:92-100: error decreases over a short train call."""
    rng = numpy.random.default_rng(1)
    x = rng.random((60, 2)) * 2 - 1
    y = numpy.stack([numpy.sin(2 * x[:, 0]), x[:, 0] * x[:, 1]], axis=1)
    random.seed(3)
    numpy.random.seed(3)
    model = L.RBF(2, 6, 2, scale_by_similarity=True)
    model.logging = False
    before = L.validation.get_error(model, x, y)
    model.train(x, y, iterations=12)
    assert L.validation.get_error(model, x, y) < before


def test_stochastic_train_runs_on_device(L, golden_traces):
    """Model.stochastic_train + select_sample (base.py:39-135): mini-batches re-uploaded per outer iteration."""
    t = golden_traces
    random.seed(0)
    numpy.random.seed(0)
    model = L.MLP((4, 6, 3), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError())
    model.logging = False
    before = L.validation.get_error(model, t["iris_x"], t["iris_y"])
    model.stochastic_train(t["iris_x"], t["iris_y"], max_iterations=6, train_kwargs={'iterations': 10})
    assert L.validation.get_error(model, t["iris_x"], t["iris_y"]) < before
