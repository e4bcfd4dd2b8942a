"""flowfusion_b200 -- B200-native (sm_100a) sampling and density evaluation for
Cosmo-Pop/flowfusion models.

Like the reference package (`flowfusion/__init__.py` is empty), the model classes live in the
submodules: ``flowfusion_b200.diffusion``, ``flowfusion_b200.flow``, ``flowfusion_b200.symplectic``.
They mirror the reference's classes (names, constructor arguments, attributes, ``state_dict``
keys, entry points and return shapes) and route the hot path -- PF-ODE / reverse-SDE / flow-ODE
integration and the divergence-trace log-likelihood -- to hand-written CUDA kernels through the
C ABI in ``include/ffb200.h``.  ``accelerate(obj)`` converts a live reference object.

There is no CPU implementation in this package: without ``libffb200.so`` or a CUDA device the
entry points raise.  The CPU oracle used by the tests lives in ``oracle/`` and is never
imported from here.
"""
from __future__ import annotations

from . import _lib

__all__ = ["accelerate", "build", "diffusion", "flow", "symplectic", "dist"]
__version__ = "0.1.0"


def build(force=False, verbose=False):
    """Compile ``libffb200.so`` in-tree (nvcc, sm_100a)."""
    return _lib.build(force=force, verbose=verbose)


def __getattr__(name):
    if name in ("diffusion", "flow", "symplectic", "dist", "engine", "solver"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)


def accelerate(obj, device="cuda"):
    """Build the flowfusion_b200 twin of a live ``flowfusion`` object and load its weights.

    Supports MLP, ScoreModel, VESDE / VPSDE / SUBVPSDE, PopulationModelDiffusion[Conditional],
    ODEFlow, ConditionalODEFlow, SymplecticMLP and SymplecticFlowModel.  Mutable flags
    (``hutch``, ``no_sigma``, ``method``, ``options``) are copied; the result is in eval mode.
    """
    import torch
    from . import diffusion as D, flow as F, symplectic as Sy

    name = type(obj).__name__

    def sde_of(s):
        n = type(s).__name__
        if n == "VESDE":
            return D.VESDE(float(s.sigma_min), float(s.sigma_max), float(s.T), float(s.epsilon))
        cls = {"VPSDE": D.VPSDE, "SUBVPSDE": D.SUBVPSDE}[n]
        return cls(s.beta_min, s.beta_max, s.T, float(s.epsilon))

    def mlp_of(m):
        units = m.architecture[1:-1]
        out = D.MLP(m.n_dimensions, m.n_conditionals, 2 * m.W.shape[0], units, activation=m.activation)
        out.load_state_dict(m.state_dict())
        return out

    def hidden_of(seq):
        lin = [l for l in seq if isinstance(l, torch.nn.Linear)]
        return [l.out_features for l in lin[:-1]]

    def act_modules_of(seq):
        """The activation modules between the Linear layers (`flow.py:41, 63-70`, `symplectic.py:25, 52-60`)."""
        acts = [m for m in seq if not isinstance(m, torch.nn.Linear)]
        if len({type(m) for m in acts}) > 1:
            raise NotImplementedError("mixed hidden-layer activations are not implemented in the CUDA kernels")
        return acts

    def act_class_of(seq):
        """Flows take the activation CLASS and instantiate it per layer with its defaults (`flow.py:67`)."""
        acts = act_modules_of(seq)
        return type(acts[0]) if acts else torch.nn.SiLU

    if name in ("VESDE", "VPSDE", "SUBVPSDE"):
        new = sde_of(obj)
    elif name == "MLP":
        new = mlp_of(obj)
    elif name == "ScoreModel":
        new = D.ScoreModel(mlp_of(obj.model), sde_of(obj.sde), no_sigma=obj.no_sigma, hutchinson=obj.hutch,
                           hutchpp=obj.hutchpp, hpp_rank=obj.hpp_rank, hpp_vecs=obj.hpp_vector,
                           xtrace=obj.xtrace, xt_vecs=obj.xt_vector)
    elif name == "PopulationModelDiffusion":
        new = D.PopulationModelDiffusion(mlp_of(obj.model), sde_of(obj.sde), obj.shift.clone(), obj.scale.clone(),
                                         method=obj.method, no_sigma=obj.score_model.no_sigma,
                                         hutchinson=obj.score_model.hutch, options=obj.options)
    elif name == "PopulationModelDiffusionConditional":
        new = D.PopulationModelDiffusionConditional(
            mlp_of(obj.model), sde_of(obj.sde), obj.shift.clone(), obj.scale.clone(), obj.conditional_shift.clone(),
            obj.conditional_scale.clone(), no_sigma=obj.score_model.no_sigma, method=obj.method, options=obj.options)
    elif name == "ODEFlow":
        new = F.ODEFlow(obj.target_dimension, hidden_of(obj.layers), activation=act_class_of(obj.layers))
        new.load_state_dict(obj.state_dict())
    elif name == "ConditionalODEFlow":
        new = F.ConditionalODEFlow(obj.target_dimension, obj.conditional_dimension, hidden_of(obj.layers),
                                   activation=act_class_of(obj.layers))
        new.load_state_dict(obj.state_dict())
    elif name == "SymplecticMLP":
        lin = [l for l in obj.mlp_q_dynamics if isinstance(l, torch.nn.Linear)]
        emb = 2 * obj.W.shape[0]
        D_ = lin[-1].out_features
        acts = act_modules_of(list(obj.mlp_q_dynamics) + list(obj.mlp_p_dynamics))
        import copy
        new = Sy.SymplecticMLP(D_, lin[0].in_features - D_ - emb, emb, hidden_of(obj.mlp_q_dynamics),
                               activation=copy.deepcopy(acts[0]) if acts else torch.nn.SiLU())   # an INSTANCE, `symplectic.py:25`
        new.load_state_dict(obj.state_dict())
    elif name == "SymplecticFlowModel":
        new = Sy.SymplecticFlowModel(accelerate(obj.model, device="cpu"), obj.shift.clone(), obj.scale.clone(),
                                     obj.conditional_shift.clone(), obj.conditional_scale.clone())
    else:
        raise TypeError(f"don't know how to accelerate a {name}")
    return new.to(device).eval()
