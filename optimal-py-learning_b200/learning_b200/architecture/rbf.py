"""Radial basis function network through the same device path as the MLP.

Reference: learning/architecture/rbf.py:36-231.  ``activate`` = cluster distances (SOM.activate) -> Gaussian
similarity -> optional scaling by the row sum -> linear output layer; the trainable part is that output layer,
whose flat parameter vector ``[bias ; W.ravel()]`` (rbf.py:229-231) is the MLP layout of a ``(clusters, outputs)``
single-layer network.  So the objective / Jacobian (rbf.py:191-225) is the MLP kernel schedule fed with the
similarity matrix, which ``rbf_similarity_kernel`` produces on the device once per dataset (or after every
clustering step when ``cluster_incrementally``); only the similarity of a public ``activate`` call is brought
back to the host (``_similarity_tensor``, as the reference exposes it).
"""
import functools
import operator

import numpy
import torch

from .. import _device as dev
from .. import error as error_mod
from ..error import MeanSquaredError
from ..optimize import Problem, make_optimizer
from .mlp import _KernelModel
from .som import SOM

INITIAL_WEIGHTS_RANGE = 0.25
LINEAR = 0


class RBF(_KernelModel):
    """Radial Basis Function network (rbf.py:36-56 for the arguments)."""

    def __init__(self, attributes, num_clusters, num_outputs, optimizer=None, error_func=None,
                 jacobian_norm_break=1e-10, variance=None, scale_by_similarity=True, clustering_model=None,
                 cluster_incrementally=False):
        super(RBF, self).__init__()
        self._init_kernel_state('f64')
        self._cluster_incrementally = cluster_incrementally
        if clustering_model is None:
            clustering_model = SOM(attributes, num_clusters, move_rate=0.1, neighborhood=2, neighbor_move_rate=1.0)
            clustering_model.logging = False
        if not isinstance(clustering_model, SOM):
            raise TypeError('clustering_model must be a learning_b200 SOM: the similarity kernel reads its neuron '
                            'weights (no CPU fallback for other clustering models)')
        self._clustering_model = clustering_model
        if variance is None:
            variance = 4.0 / num_clusters
        self._variance = variance

        self._shape = (num_clusters, num_outputs)
        self._lowered_transfers = [(LINEAR, 0.0)]
        self._weight_matrix = self._random_weight_matrix(self._shape)
        self._bias_vec = self._random_weight_matrix(self._shape[1])
        if optimizer is None:
            optimizer = make_optimizer(functools.reduce(operator.mul, self._weight_matrix.shape))
        self._optimizer = optimizer
        if error_func is None:
            error_func = MeanSquaredError()
        self._error_func = error_func
        self._lowered_error = error_mod.lower(error_func)
        self._jacobian_norm_break = jacobian_norm_break
        self._scale_by_similarity = scale_by_similarity
        self._similarity_tensor = None
        self._features = None          # (key, similarity matrix on the device, targets on the device)

    def reset(self):
        super(RBF, self).reset()
        self._clustering_model.reset()
        self._optimizer.reset()
        self._weight_matrix = self._random_weight_matrix(self._weight_matrix.shape)
        self._bias_vec = self._random_weight_matrix(self._shape[1])
        self._similarity_tensor = None
        self._features = None

    def _random_weight_matrix(self, shape):
        return (2 * numpy.random.random(shape) - 1) * INITIAL_WEIGHTS_RANGE

    # ------------------------------------------------------------------ forward
    def _similarity_device(self, x_dev):
        return self._clustering_model.distances_device(x_dev, 2 if self._scale_by_similarity else 1, self._variance)

    def activate(self, input_tensor):
        """Model outputs for one pattern or a matrix of patterns (rbf.py:133-151)."""
        if not isinstance(input_tensor, numpy.ndarray):
            input_tensor = numpy.array(input_tensor, dtype=numpy.float64)
        vector = input_tensor.ndim == 1
        x = numpy.ascontiguousarray(input_tensor, dtype=numpy.float64).reshape(-1, input_tensor.shape[-1])
        if x.shape[-1] != self._clustering_model._size[1]:
            raise ValueError('input_tensor attributes == %s, expected %s' % (x.shape[-1], self._clustering_model._size[1]))
        sim = self._similarity_device(dev.to_device(x))
        host_sim = dev.to_host(sim)
        self._similarity_tensor = host_sim[0] if vector else host_sim
        out = self._forward_host(_flatten_weights(self._weight_matrix, self._bias_vec), host_sim)
        return out[0] if vector else out

    # ------------------------------------------------------------------ training
    def _training_features(self, input_matrix, target_matrix):
        """(similarity matrix, targets) of this rank's rows on the device; recomputed when the data or the cluster
        centres change."""
        x = numpy.asarray(input_matrix, dtype=numpy.float64)
        y = numpy.asarray(target_matrix, dtype=numpy.float64)
        if x.ndim != 2 or y.ndim != 2 or x.shape[0] != y.shape[0]:
            raise ValueError('input_matrix and target_matrix must be matrices with equal row counts')
        if y.shape[-1] != self._shape[-1]:
            raise ValueError('target attributes == %s, expected %s' % (y.shape[-1], self._shape[-1]))
        # cached only inside a residency scope (train / train_step) on these very objects and for these cluster centres
        scope = self._residency_scope
        in_scope = scope is not None and scope[0] is input_matrix and scope[1] is target_matrix
        centres = self._clustering_model._weights.tobytes()
        cached = self._features
        if in_scope and cached is not None and cached[0] is input_matrix and cached[1] is target_matrix and cached[2] == centres:
            return cached[3], cached[4]
        first, count, _ = self._shard(x.shape[0])
        x_dev = dev.to_device(numpy.ascontiguousarray(x[first:first + count]))
        features = (input_matrix, target_matrix, centres, self._similarity_device(x_dev),
                    dev.to_device(numpy.ascontiguousarray(y[first:first + count])))
        self._features = features if in_scope else None
        return features[3], features[4]

    def _enter_scope(self, input_matrix, target_matrix):
        self._features = None          # the similarity matrix is rebuilt on first use inside the scope

    def train_step(self, input_matrix, target_matrix):
        """One optimizer step on the output layer (rbf.py:153-177)."""
        with self.resident(input_matrix, target_matrix):
            return self._train_step_resident(input_matrix, target_matrix)

    def _train_step_resident(self, input_matrix, target_matrix):
        if self._cluster_incrementally:
            self._clustering_model.train_step(input_matrix, target_matrix)
        sim, tgt = self._training_features(input_matrix, target_matrix)
        ws = self._make_resident(sim, tgt)
        problem = Problem(obj_func=lambda xk: self._get_obj(xk, input_matrix, target_matrix),
                          obj_jac_func=lambda xk: self._get_obj_jac(xk, input_matrix, target_matrix))
        problem.ray_obj_jac = lambda xk, alpha, direction: self._probe(ws, xk, alpha, direction, True)[:2]
        problem.ray_obj = lambda xk, alpha, direction: self._probe(ws, xk, alpha, direction, False)[0]
        error, flat = self._optimizer.next(problem, dev.to_device(_flatten_weights(self._weight_matrix, self._bias_vec)))
        self._bias_vec, self._weight_matrix = _unflatten_weights(dev.to_host(flat), self._shape)
        jacobian = self._optimizer.jacobian
        if jacobian is None:
            self.converged = False
        else:
            norm = dev.norm(jacobian) if dev.is_device_vector(jacobian) else float(numpy.linalg.norm(jacobian))
            self.converged = norm < self._jacobian_norm_break
        return error

    def _pre_train(self, input_matrix, target_matrix):
        if not self._cluster_incrementally:
            self._clustering_model.train(input_matrix, target_matrix)      # cluster the input space (rbf.py:184-186)
        super(RBF, self)._pre_train(input_matrix, target_matrix)

    # ------------------------------------------------------------------ objective / gradient (rbf.py:191-225)
    def _get_obj(self, parameter_vec, input_matrix, target_matrix):
        sim, tgt = self._training_features(input_matrix, target_matrix)
        return self._objective(parameter_vec, sim, tgt, False)[0]

    def _get_obj_jac(self, parameter_vec, input_matrix, target_matrix):
        sim, tgt = self._training_features(input_matrix, target_matrix)
        return self._objective(parameter_vec, sim, tgt, True)

    def _adopt(self, params):
        self._bias_vec, self._weight_matrix = _unflatten_weights(params, self._shape)

    def __getstate__(self):
        state = super(RBF, self).__getstate__()
        state['_features'] = None
        return state


def _flatten_weights(weight_matrix, bias_vec):
    """Flat vector of model parameters (rbf.py:229-231)."""
    return numpy.hstack([bias_vec, weight_matrix.ravel()])


def _unflatten_weights(flat_weights, shape):
    """(bias, weight matrix) views of the flat vector (rbf.py:234-236)."""
    return flat_weights[:shape[1]], flat_weights[shape[1]:].reshape(shape)
