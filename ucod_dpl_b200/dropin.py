"""Drop-in switch: make the reference's import paths resolve to this package.

The reference has no plugin/FFI layer — its boundary is the Python module API (`from models.uscod import baseline`,
`from data.utils.feature_extractor import backbone`, `from engine.config import CfgNode`, ... see SURVEY.md §8b).
`install()` registers this package's modules in `sys.modules` under those names, so the reference's launch scripts
(scripts/eval.py, scripts/LTeval.py, generate_pseudo_label.py) import the B200-native implementation unchanged.
Importing this module also imports every hot-path module, which populates the registries of `engine.registry`.
"""
from __future__ import annotations

import importlib
import sys

ALIASES = {
    "engine": "ucod_dpl_b200.engine",
    "engine.config": "ucod_dpl_b200.engine.config",
    "engine.config.config": "ucod_dpl_b200.engine.config.config",
    "engine.registry": "ucod_dpl_b200.engine.registry",
    "engine.registry.registry": "ucod_dpl_b200.engine.registry.registry",
    "engine.registry.root": "ucod_dpl_b200.engine.registry.root",
    "engine.utils": "ucod_dpl_b200.engine.utils",
    "engine.utils.fileio": "ucod_dpl_b200.engine.utils.fileio",
    "engine.utils.metrics": "ucod_dpl_b200.engine.utils.metrics",
    "engine.utils.metrics.metric": "ucod_dpl_b200.engine.utils.metrics.metric",
    "engine.runner": "ucod_dpl_b200.engine.runner",
    "engine.runner.loop_UCOD_DPL": "ucod_dpl_b200.engine.runner.loop_UCOD_DPL",
    "engine.runner.loop_CORAL": "ucod_dpl_b200.engine.runner.loop_CORAL",
    "models": "ucod_dpl_b200.models",
    "models.uscod": "ucod_dpl_b200.models.uscod",
    "models.discriminator": "ucod_dpl_b200.models.discriminator",
    "models.UDLR": "ucod_dpl_b200.models.UDLR",
    "models.modules": "ucod_dpl_b200.models.modules",
    "models.modules.DBA": "ucod_dpl_b200.models.modules.DBA",
    "models.modules.ASR": "ucod_dpl_b200.models.modules.refiner",
    "models.modules.HRE": "ucod_dpl_b200.models.modules.refiner",
    "models.modules.CSF": "ucod_dpl_b200.models.modules.refiner",
    "models.modules.GE_pix_level": "ucod_dpl_b200.models.modules.refiner",
    "models.modules.mlp": "ucod_dpl_b200.models.modules.refiner",
    "data": "ucod_dpl_b200.data",
    "data.utils": "ucod_dpl_b200.data.utils",
    "data.utils.feature_extractor": "ucod_dpl_b200.data.utils.feature_extractor",
    "data.utils.found_bkg_mask": "ucod_dpl_b200.data.utils.found_bkg_mask",
    "data.datasets": "ucod_dpl_b200.data.datasets",
    "data.datasets.transforms": "ucod_dpl_b200.data.datasets.transforms",
    "data.datasets.cache_manager": "ucod_dpl_b200.data.datasets.cache_manager",
    "data.datasets.base_dataset": "ucod_dpl_b200.data.datasets.base_dataset",
    "data.datasets.uscod_dataset": "ucod_dpl_b200.data.datasets.base_dataset",
    "data.datasets.lr_dataset": "ucod_dpl_b200.data.datasets.lr_dataset",
    "data.datasets.dataloader_utils": "ucod_dpl_b200.data.datasets.dataloader_utils",
    "engine.utils.save_image": "ucod_dpl_b200.engine.utils.save_image",
    "scripts": "ucod_dpl_b200.scripts",
    "scripts.args": "ucod_dpl_b200.scripts.args",
    "scripts.eval": "ucod_dpl_b200.scripts.eval",
    "scripts.train": "ucod_dpl_b200.scripts.train",
    "scripts.LTeval": "ucod_dpl_b200.scripts.LTeval",
    "generate_pseudo_label": "ucod_dpl_b200.generate_pseudo_label",
}

_loaded = {name: importlib.import_module(target) for name, target in ALIASES.items()}


def install(force: bool = False) -> None:
    """Alias the reference's module names to this package (refuses to shadow foreign modules unless `force`)."""
    for name, mod in _loaded.items():
        cur = sys.modules.get(name)
        if cur is not None and cur is not mod and not force:
            raise ImportError(f"module '{name}' is already imported from {getattr(cur, '__file__', '?')}; "
                              "call install(force=True) to replace it")
        sys.modules[name] = mod


def uninstall() -> None:
    for name, mod in _loaded.items():
        if sys.modules.get(name) is mod:
            del sys.modules[name]
