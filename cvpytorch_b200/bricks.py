"""Brick-level drop-in: the reference builds its conv bundles through registries --
``CONV_LAYERS.register_module(name, module=cls)`` (src/models/bricks/registry.py:4-9, src/utils/registry.py:293-346) and
``ConvModule(conv_cfg=dict(type=...), norm_cfg=..., act_cfg=...)`` (src/models/bricks/conv_module.py:74-89).

A maintainer who wants single bricks of an otherwise unchanged reference model on the B200 path adds, once at import time::

    from src.models.bricks import registry            # the reference's registries
    import cvpytorch_b200.bricks as b200
    b200.register(registry)                            # CONV_LAYERS['B200Conv2d'], PLUGIN_LAYERS['B200ConvModule']

and then either builds whole bundles with ``build_plugin_layer(dict(type='B200ConvModule', in_channels=..., out_channels=..., ...))``
or swaps the class the reference's model code instantiates (``conv_module.ConvModule = b200.ConvModule``).  ``B200ConvModule`` is
:class:`cvpytorch_b200.modules.ConvModule`: same constructor, same ``conv.weight`` / ``bn.*`` state_dict keys, ``forward(x)`` NCHW fp32
in / out, conv + folded eval-mode BN + activation as ONE tcgen05 kernel.  ``B200Conv2d`` is the marker the reference's
``build_conv_layer`` resolves for ``conv_cfg=dict(type='B200Conv2d')``: an ``nn.Conv2d`` subclass holding the parameters (so
checkpoints keep their keys); the arithmetic of a bundle always runs in the fused kernel of the enclosing B200 ConvModule.
"""
import torch.nn as nn

from .modules import ConvModule


class B200Conv2d(nn.Conv2d):
    """Parameter holder registered under CONV_LAYERS['B200Conv2d'] (constructor of nn.Conv2d)."""

    def forward(self, x):  # pragma: no cover - guarded path
        raise RuntimeError('B200Conv2d is executed by its enclosing cvpytorch_b200.modules.ConvModule (conv + BN + activation fused into one '
                           'tcgen05 kernel); wrap it in ConvModule(conv_cfg=dict(type="B200Conv2d"), ...) and call the bundle')


B200ConvModule = ConvModule


def register(registry_module):
    """registry_module: the reference's ``src.models.bricks.registry`` (anything with CONV_LAYERS / PLUGIN_LAYERS objects that offer
    ``register_module(name=, module=)``).  Idempotent."""
    for reg_name, name, cls in (('CONV_LAYERS', 'B200Conv2d', B200Conv2d), ('PLUGIN_LAYERS', 'B200ConvModule', B200ConvModule)):
        reg = getattr(registry_module, reg_name)
        try:
            reg.register_module(name=name, module=cls)
        except KeyError:  # already registered
            pass
    return registry_module
