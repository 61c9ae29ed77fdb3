"""The two validation helpers the hot path's callers use, batched.

The reference evaluates ``model.activate`` one pattern at a time in a Python loop
(learning/validation.py:235-254); here the whole matrix goes through one batched forward
(the same kernels as training), which is the SURVEY.md section 8(f) rank-1 widening.
"""
import numpy

from .error import MeanSquaredError


def get_error(model, input_matrix, target_matrix, error_func=None):
    """Mean over patterns of error_func(model(x), target) (validation.py:235-244).

    The reference averages per-pattern errors; for MSE and cross entropy the mean of the
    per-row errors equals the error of the whole matrix, which is what the row kernel returns.
    """
    if error_func is None:
        error_func = MeanSquaredError()
    outputs = model.activate(numpy.asarray(input_matrix, dtype=float))
    return error_func(outputs, numpy.asarray(target_matrix, dtype=float))


def _get_classes(matrix):
    """Class of every row: the label itself for a column matrix, else the arg-max (validation.py:257-271)."""
    matrix = numpy.asarray(matrix)
    if matrix.ndim == 1:
        return matrix
    if matrix.shape[1] == 1:
        return matrix.ravel()
    return numpy.argmax(matrix, axis=1)


def get_accuracy(model, input_matrix, target_matrix):
    """Fraction of patterns whose class matches the target's (validation.py:247-254): a column matrix holds the class
    labels themselves, a wider one is one-hot / scores and is arg-maxed (``_get_classes``)."""
    outputs = numpy.asarray(model.activate(numpy.asarray(input_matrix, dtype=float)))
    hits = _get_classes(outputs) == _get_classes(numpy.asarray(target_matrix))
    return float(numpy.mean(hits))


def make_train_test_sets(input_matrix, label_matrix, train_per_class):
    """((training_inputs, training_labels), (testing_inputs, testing_labels)): the first ``train_per_class`` patterns of
    every distinct label row go to the training set, the rest to the testing set, order preserved
    (validation.py:342-381; host bookkeeping of the README example, examples/readme_example.py:33-36)."""
    inputs, labels = numpy.asarray(input_matrix), numpy.asarray(label_matrix)
    seen = {}
    to_train = numpy.zeros(len(labels), dtype=bool)
    for row, label in enumerate(labels):
        key = tuple(label.tolist()) if label.ndim else (label.item(),)
        seen[key] = seen.get(key, 0) + 1
        to_train[row] = seen[key] <= train_per_class
    if to_train.all():
        raise ValueError('train_per_class too high, no testing set')
    return (inputs[to_train], labels[to_train]), (inputs[~to_train], labels[~to_train])
