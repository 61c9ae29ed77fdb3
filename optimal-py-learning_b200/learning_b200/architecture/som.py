"""Self-organizing map on the device (the default clustering model of the RBF network).

Reference: learning/architecture/som.py:33-113.  Same constructor, attributes (``move_rate``,
``neighborhood``, ``neighbor_move_rate``, ``_weights``, ``_distances``) and seeded weight draw.  The
competition/move loop that the reference runs pattern by pattern in Python (base.py:289-331 ->
``_train_increment`` -> ``_move_neurons``) is sequential by construction; here one ``train_step`` is one
launch of ``som_train_kernel`` (a single resident CTA walking the patterns in order with the neuron
weights in shared memory), and ``activate`` is the batched distance kernel.
"""
import ctypes

import numpy
import torch

from .. import _device as dev
from .. import _native as nat
from ..base import Model


def _gaussian(x, variance=1.0):
    """calculate.gaussian (calculate.py:101-102) for the handful of neighbour move rates (host scalars)."""
    return numpy.exp(-(x**2 / variance))


class SOM(Model):
    def __init__(self, attributes, neurons, move_rate=0.1, neighborhood=2, neighbor_move_rate=1.0,
                 initial_weights_range=1.0):
        super(SOM, self).__init__()
        self.move_rate = move_rate
        self.neighborhood = neighborhood
        self.neighbor_move_rate = neighbor_move_rate
        self.initial_weights_range = initial_weights_range
        self._size = (neurons, attributes)
        self._weights = numpy.zeros(self._size)
        self._distances = numpy.zeros(neurons)
        self._resident = None          # (training matrix object, its device copy) inside a train() call
        self._train_scope_input = None
        self.reset()

    def reset(self):
        """Randomize weights between -range and range (som.py:51-59)."""
        super(SOM, self).reset()
        self._weights = (2 * numpy.random.random(self._size) - 1) * self.initial_weights_range
        self._distances = numpy.zeros(self._size)

    # ------------------------------------------------------------------ forward
    def _check_input(self, input_tensor):
        if not isinstance(input_tensor, numpy.ndarray):
            input_tensor = numpy.array(input_tensor)
        if input_tensor.ndim not in (1, 2):
            raise ValueError('Invalid shape of input_tensor.')
        if input_tensor.shape[-1] != self._size[1]:
            raise ValueError('input_tensor attributes == %s, expected %s' % (input_tensor.shape[-1], self._size[1]))
        return input_tensor

    def distances_device(self, x_dev, mode=0, variance=1.0):
        """rows x neurons distances (mode 0) / Gaussian similarities (1) / scaled similarities (2) of a device matrix."""
        out = torch.empty((x_dev.shape[0], self._size[0]), dtype=torch.float64, device=x_dev.device)
        centers = dev.to_device(self._weights)
        nat.check(nat.lib.opl_rbf_similarity_device(dev.ptr(x_dev), x_dev.shape[0], self._size[1], dev.ptr(centers),
                                                    self._size[0], int(mode), float(variance), dev.ptr(out),
                                                    dev.stream_ptr()))
        return out

    def activate(self, input_tensor):
        """Distance of every neuron to the pattern(s) (som.py:61-85)."""
        x = self._check_input(input_tensor)
        out = dev.to_host(self.distances_device(dev.to_device(x.reshape(-1, self._size[1]))))
        self._distances = out[0] if x.ndim == 1 else out
        return self._distances

    # ------------------------------------------------------------------ training
    def _rates(self):
        """Final move rate by distance from the winner (som.py:108-111)."""
        return numpy.array([float(_gaussian(float(d), self.neighbor_move_rate)) * self.move_rate
                            for d in range(int(self.neighborhood) + 1)], dtype=numpy.float64)

    def _train_on_device(self, x_dev, passes=1):
        w = dev.to_device(self._weights)
        rates = dev.to_device(self._rates())
        nat.check(nat.lib.opl_som_train_device(dev.ptr(x_dev), x_dev.shape[0], self._size[1], dev.ptr(w), self._size[0],
                                               int(passes), dev.ptr(rates), int(self.neighborhood), dev.stream_ptr()))
        self._weights = dev.to_host(w)

    def _device_matrix(self, input_matrix):
        """Device copy of the training matrix: uploaded once per train() call (kept between its train_steps for the very
        object train() was given), on every call otherwise -- never guessed from the contents."""
        cached = self._resident
        if cached is not None and cached[0] is input_matrix:
            return cached[1]
        x = self._check_input(input_matrix)
        x_dev = dev.to_device(x.reshape(-1, self._size[1]))
        if self._train_scope_input is input_matrix:
            self._resident = (input_matrix, x_dev)
        return x_dev

    def _pre_train(self, input_matrix, target_matrix):
        self._train_scope_input = input_matrix
        self._resident = None

    def train_step(self, input_matrix, target_matrix):
        """One pass over the patterns in order (base.py:289-331 with SOM._train_increment).  Returns None, as the
        reference's SOM reports no error."""
        if self._post_pattern_callback:
            return super(SOM, self).train_step(input_matrix, target_matrix)     # per-pattern path keeps the callback
        self._train_on_device(self._device_matrix(input_matrix))
        return None

    def _train_increment(self, input_vec, target_vec):
        """One pattern (som.py:87-98)."""
        x = self._check_input(input_vec)
        self._train_on_device(dev.to_device(x.reshape(1, -1)))

    def _post_train(self, input_matrix, target_matrix):
        self._train_scope_input = None
        self._resident = None

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_resident'] = None
        state['_train_scope_input'] = None
        return state
