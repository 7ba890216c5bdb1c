"""Configuration objects with the attribute layout the reference reads from ``src/configs/config.yaml`` through
OmegaConf (``model.use_lb``, ``model.use_dgr``, ``model.loss_function.*``; config.yaml:22-42, main.py:66-79).

`omegaconf` is not a dependency: any object exposing these attributes works (an OmegaConf node does too)."""
import types


class Cfg(types.SimpleNamespace):
    pass


def model_cfg(use_lb=False, use_dgr=False):
    """the `cfg` argument of MidasNet_small / MidasNetSemantics (midas_net_custom.py:63-64); config.yaml keeps both off"""
    return Cfg(use_lb=use_lb, use_dgr=use_dgr)


def loss_config(si=1.0, silog=0.0, vf=0.85, grad=0.0, edge=0.0):
    """the `config` argument of combined_loss (main.py:51-89); defaults = config.yaml:34-42 (1 / 0 / 0 / 0)"""
    return Cfg(model=Cfg(loss_function=Cfg(si_loss_alpha=si, silog_loss=Cfg(alpha=silog, variance_focus=vf),
                                           grad_loss_alpha=grad, edge_loss_alpha=edge)))
