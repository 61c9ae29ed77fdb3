"""Model base class: the training loop that drives ``train_step``.

Host-side control flow only (no arithmetic): iteration / retry / stopping rules follow
learning/base.py:77-373 so that a model trained through this package stops at the same
iteration, for the same reason, as the reference.
"""
import math
import pickle
import random

from . import validation


def _sample_rows(count, size=None):
    """Row numbers of select_sample: the same ``random`` calls, in the same order, as base.py:39-50."""
    if size is None:
        size = _selection_size_heuristic(count)
    return random.sample(range(count), size)


def _random_rows(count, size=None):
    """Row numbers of select_random (base.py:53-64)."""
    if size is None:
        size = _selection_size_heuristic(count)
    return [random.randint(0, count - 1) for _ in range(size)]


def select_sample(input_matrix, target_matrix, size=None):
    """Random rows without replacement, random order (base.py:39-50)."""
    rows = _sample_rows(input_matrix.shape[0], size)
    return input_matrix[rows], target_matrix[rows]


def select_random(input_matrix, target_matrix, size=None):
    """Random rows with replacement (base.py:53-64)."""
    rows = _random_rows(input_matrix.shape[0], size)
    return input_matrix[rows], target_matrix[rows]


# selection functions whose row choice can be replayed on a device-resident dataset
ROW_SELECTORS = {select_sample: _sample_rows, select_random: _random_rows}


def _selection_size_heuristic(num_samples):
    """Mini-batch size growing with log2 of the dataset size (base.py:67-74)."""
    return min(num_samples, int(20.0 * (math.log(num_samples + 10, 2) - math.log(10.0) + 1.0)))


class Model(object):
    """A supervised learning model (base.py:77)."""

    def __init__(self):
        self._post_pattern_callback = None
        self.logging = True
        self.iteration = 0
        self.converged = False

    def _reset_bookkeeping(self):
        self.iteration = 0

    def reset(self):
        self.iteration = 0
        self.converged = False

    def activate(self, input_vec):
        raise NotImplementedError()

    # ------------------------------------------------------------------ training loops
    def stochastic_train(self, input_matrix, target_matrix, max_iterations=100, error_break=0.002,
                         pattern_selection_func=select_sample, train_kwargs={'iterations': 100}):
        """Repeated ``train`` on random subsets (base.py:100-135).  Models with a device-side dataset
        (``_select_patterns``) keep the whole dataset resident and select each mini-batch there."""
        select = getattr(self, '_select_patterns', None)
        for iteration in range(1, max_iterations + 1):
            if select is not None:
                batch = select(pattern_selection_func, input_matrix, target_matrix)
            else:
                batch = pattern_selection_func(input_matrix, target_matrix)
            train_error = self.train(*batch, **train_kwargs)
            if self.converged:
                if (train_error <= error_break
                        and validation.get_error(self, input_matrix, target_matrix) <= error_break):
                    return train_error
        self.iteration = iteration
        return train_error

    def train(self, input_matrix, target_matrix, iterations=1000, retries=0, error_break=0.002,
              error_stagnant_distance=5, error_stagnant_threshold=0.00001, error_improve_iters=20,
              post_pattern_callback=None):
        """Train on the given dataset (base.py:137-182); returns the last training error."""
        self.converged = False
        self._pre_train(input_matrix, target_matrix)
        try:
            train_error = self._train(input_matrix, target_matrix, iterations, retries, error_break,
                                      error_stagnant_distance, error_stagnant_threshold, error_improve_iters,
                                      post_pattern_callback)
        finally:
            self._post_train(input_matrix, target_matrix)
        return train_error

    def _train(self, input_matrix, target_matrix, iterations, retries, error_break, error_stagnant_distance,
               error_stagnant_threshold, error_improve_iters, post_pattern_callback):
        """Attempts with reset in between, keeping the best serialized attempt (base.py:184-228)."""
        self._reset_bookkeeping()
        self._post_pattern_callback = post_pattern_callback
        best_error, best_blob = float('inf'), None
        attempt_error = None
        for attempt in range(retries + 1):
            attempt_error = self._train_attempt(input_matrix, target_matrix, iterations, error_break,
                                                error_stagnant_distance, error_stagnant_threshold,
                                                error_improve_iters, post_pattern_callback)
            if self.converged or retries == 0:
                return attempt_error
            if attempt >= retries:
                if not attempt_error < best_error:
                    self.__dict__ = self.unserialize(best_blob).__dict__
                return attempt_error
            if attempt_error < best_error:
                best_error, best_blob = attempt_error, self.serialize()
            self.reset()
        return attempt_error

    def _train_attempt(self, input_matrix, target_matrix, iterations, error_break, error_stagnant_distance,
                       error_stagnant_threshold, error_improve_iters, post_pattern_callback):
        """One run of up to ``iterations`` train_steps with the stopping rules of base.py:230-287."""
        recent = [1e10] * error_stagnant_distance
        best_error = float('inf')
        since_improvement = 0
        error = None
        for self.iteration in range(1, iterations + 1):
            error = self.train_step(input_matrix, target_matrix)
            if self.logging:
                print("Iteration {}, Error: {}".format(self.iteration, error))
            if self.converged:
                return error
            if error is None:
                continue
            if error <= error_break:
                self.converged = True
                return error
            if self.iteration == iterations:
                return error
            if all(abs(previous - error) <= error_stagnant_threshold for previous in recent):
                return error           # stagnated
            recent.append(error)
            recent.pop(0)
            if error < best_error:
                best_error, since_improvement = error, 0
            else:
                since_improvement += 1
            if since_improvement >= error_improve_iters:
                return error           # no improvement
        return error

    def train_step(self, input_matrix, target_matrix):
        """Default: per-pattern ``_train_increment`` and mean error (base.py:289-331)."""
        total = 0.0
        for input_vec, target_vec in zip(input_matrix, target_matrix):
            step_error = self._train_increment(input_vec, target_vec)
            if step_error is not None and not isinstance(step_error, (int, float)):
                raise TypeError('%s._train_increment must return an error number or None' % type(self))
            if self._post_pattern_callback:
                self._post_pattern_callback(self, input_vec, target_vec)
            if total is not None and step_error is not None:
                total += step_error
            else:
                total = None
        return None if total is None else total / len(input_matrix)

    def _train_increment(self, input_vec, target_vec):
        raise NotImplementedError()

    def _pre_train(self, input_matrix, target_matrix):
        pass

    def _post_train(self, input_matrix, target_matrix):
        pass

    # ------------------------------------------------------------------ checkpointing
    def serialize(self):
        """Pickle (protocol 2) of the model; device state is converted to host arrays by the
        subclasses' ``__getstate__`` (base.py:350-360)."""
        return pickle.dumps(self, protocol=2)

    @classmethod
    def unserialize(cls, serialized_model):
        model = pickle.loads(serialized_model)
        if type(model) != cls:
            raise ValueError('serialized_model does not match this class')
        return model

    def print_results(self, input_matrix, target_matrix):
        """Print corresponding inputs and outputs from a dataset (base.py:378-381): the same lines as the reference's
        per-pattern loop, from ONE batched activate when the model takes a matrix of patterns."""
        outputs = None
        try:
            import numpy
            batch = self.activate(numpy.asarray(input_matrix))
            if len(batch) == len(input_matrix):
                outputs = batch
        except Exception:
            outputs = None                    # a model that only takes single patterns: the reference's loop
        for row, (inp_vec, tar_vec) in enumerate(zip(input_matrix, target_matrix)):
            out_vec = self.activate(inp_vec) if outputs is None else outputs[row]
            print(inp_vec, '->', out_vec, '(%s)' % tar_vec)
