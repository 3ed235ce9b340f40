"""openeat_b200 -- B200-native (sm_100a) acoustic front-end for OpenEAT / wenet-style recipes.

Drop-in replacements for the reference's front-end callables (same names, arguments and
error behaviour), backed by hand-written CUDA kernels behind a thin C ABI:

    openeat.dataset.dataset._extract_feature / audio_collate_func  -> openeat_b200.dataset
    openeat.dataset.audio_processor._speed_generator / _speed_perturb -> openeat_b200.audio_processor
    openeat.dataset.feature_processor._normalization / _spec_augmentation / _spec_substitute
                                                                   -> openeat_b200.feature_processor
    openeat.modules.cmvn.GlobalCMVN, openeat.utils.cmvn.load_cmvn  -> openeat_b200.cmvn
    (absent in the reference) compute_cmvn_stats                    -> openeat_b200.cmvn

There is no CPU fallback: importing is cheap, but any compute call raises
``FrontendError`` when the CUDA library has not been built or no GPU is present.
"""
from ._lib import FrontendError, build  # noqa: F401

__all__ = ['FrontendError', 'build']
