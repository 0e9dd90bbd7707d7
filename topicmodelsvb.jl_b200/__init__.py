"""B200-native coordinate-ascent VB engine behind the gpuLDA / gpuCTM / gpuCTPF + train! surface of
ericproffitt/TopicModelsVB.jl.  The device code is topicmodelsvb.jl_b200/csrc (sm_100a CUDA, C ABI in
include/tmvb.h); this package is the host-side mirror of the reference's Julia API for that path.
"""
from . import _lib, synth  # noqa: F401
from ._lib import TopicModelError, build  # noqa: F401
from .corpus import Corpus, CorpusError, Document, DocumentError  # noqa: F401
from .gpu_lda import check_elbo, check_model, gpuLDA, train  # noqa: F401
