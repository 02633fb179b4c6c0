"""`neuron_models.simulated` and the public Gabor-bank generators it is assembled from.

Mirror of laminr/neuron_models/simulated_models.py:74-139 (`neuron2_generator`), :243-274
(`phase_invariant_neuron_generator`) and :380-394 (`simulated`).  Generators return `(model, None)`: the second
element of the reference's tuple is a plotting grid (visualisation only, out of scope)."""
import numpy as np

from .utils.simulated_model import (ArbitraryMultiNeuronModel, ArbitraryNeuronModel, gabor_geometry, generate_gabor_filter,
                                    random_transformation_matrices)


def phase_invariant_neuron_generator(loc, img_res, transformation_matrix=None, unit_norm=True, num_filters=32):
    filters = []
    # the filters of the bank share centre, widths and transform: coordinates and envelope once, one cosine per phase (same values as a
    # `generate_gabor_filter` call per phase, simulated_models.py:243-274)
    xx, envelope = gabor_geometry(theta=0, width_x=0.14, width_y=0.14, center=loc, grid_size=img_res, transformation_matrix=transformation_matrix)
    carrier = 2 * np.pi * 2.3 * xx
    for phase in np.linspace(0, 2 * np.pi, num_filters, endpoint=False):
        gb = envelope * np.cos(carrier + phase)
        if unit_norm:
            gb = gb / np.linalg.norm(gb)
        filters.append(gb.astype(np.float32))
    return ArbitraryNeuronModel(np.stack(filters)), None


def neuron2_generator(loc, img_res, transformation_matrix=None, unit_norm=True, num_filters=30):
    sigma_x, sigma_y, sf, min_val, max_val, periods = 0.11, 0.15, 3, 0.9, 1.5, 1
    drange = max_val - min_val
    gabor_locs = np.array([[-0.09, 0.01], [-0.01, -0.05], [0.11, 0.04]]) + loc
    thetas = (-np.pi / 8, 0, np.pi / 8)
    filters = []
    for x in np.linspace(0, 2 * np.pi * periods, periods * num_filters):
        scales = ((np.sin(x) + 1) / 2 * drange + min_val, (np.sin(x + np.pi / 3) + 1) / 2 * drange + min_val,
                  (np.sin(x + np.pi / 3 * 2) + 1) / 2 * drange + min_val)
        parts = [generate_gabor_filter(theta=thetas[k], width_x=sigma_x * scales[k], width_y=sigma_y * scales[k], frequency=sf / scales[k], phase=0,
                                       center=gabor_locs[k], grid_size=img_res, transformation_matrix=transformation_matrix) for k in range(3)]
        gb = (parts[0] + parts[1] + parts[2]) / 3
        if unit_norm:
            gb = gb / np.linalg.norm(gb)
        filters.append(gb.astype(np.float32))
    return ArbitraryNeuronModel(np.stack(filters)), None


def simulated(model_type="demo1", img_res=[100, 100]):
    if model_type == "demo1":
        neuron_model1, _ = phase_invariant_neuron_generator([0.2, 0.2], img_res, num_filters=20)
        t_mat = np.array([[0.70710678, -0.70710678], [0.70710678, 0.70710678]])
        neuron_model2, _ = phase_invariant_neuron_generator([-0.2, -0.2], img_res, transformation_matrix=t_mat, num_filters=20)
        neuron_model3, _ = neuron2_generator([-0.2, 0.2], img_res, num_filters=60)
        return ArbitraryMultiNeuronModel([neuron_model1, neuron_model2, neuron_model3])
    raise ValueError(f"Model type {model_type} does not exist.")


def simulated_population(n_neurons, img_res=[100, 100], seed=0, num_filters=20, only=None):
    """Synthetic population of phase-invariant neurons assembled from the public generators (BASELINE config 3):
    np.random.seed(seed); mats = random_transformation_matrices(n); locs ~ U(-0.35, 0.35)^2 drawn after the matrices;
    neuron i uses transformation mats[i] (None for i = 0).
    `only` (optional list of population indices): build just these neurons of the `n_neurons` population -- the same random stream, so neuron k of
    the result is bit-identical to neuron only[k] of the full population (a sample of config 3 without its 819 MB bank)."""
    np.random.seed(seed)
    mats = random_transformation_matrices(n_neurons)
    locs = np.random.uniform(-0.35, 0.35, size=(n_neurons, 2))
    import torch
    idxs = list(range(n_neurons)) if only is None else [int(i) for i in only]
    packed = torch.empty(len(idxs), num_filters, int(img_res[1]), int(img_res[0]))  # meshgrid 'xy': rows = grid_size[1]
    models = []
    for k, i in enumerate(idxs):
        m = phase_invariant_neuron_generator(locs[i], img_res, transformation_matrix=None if i == 0 else mats[i], num_filters=num_filters)[0]
        packed[k].copy_(m.filters)
        m.filters = packed[k]  # the per-neuron buffer becomes a view of the packed bank: one copy of the population in host memory
        models.append(m)
    return ArbitraryMultiNeuronModel(models, packed=packed)
