"""GPU: SURVEY section 8(f) rows -- DropoutMLP, Linear / Logistic regression with penalties, batched
validation -- against golden vectors produced by the real reference (oracle/gen_golden.py part E)."""
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
