"""B200-native coordinate-ascent VB engine behind the gpuLDA / gpuCTM / gpuCTPF + train! surface of
ericproffitt/TopicModelsVB.jl.  The device code is topicmodelsvb.jl_b200/csrc (sm_100a CUDA, C ABI in
include/tmvb.h); this package is the host-side mirror of the reference's Julia API for that path.
"""
from . import _lib, synth  # noqa: F401
from ._lib import TopicModelError, build  # noqa: F401
from .corpus import Corpus, CorpusError, Document, DocumentError, readcorp  # noqa: F401
from .gpu_ctm import check_model_ctm, gpuCTM, gpufCTM, train_ctm  # noqa: F401
from .gpu_ctpf import check_model_ctpf, gpuCTPF, train_ctpf  # noqa: F401
from .gpu_flda import check_model_flda, gpufLDA, train_flda  # noqa: F401
from .gpu_lda import check_elbo, gpuLDA  # noqa: F401
from .gpu_lda import check_model as check_model_lda  # noqa: F401
from .gpu_lda import train as train_lda  # noqa: F401

from .predict import predict, topicdist  # noqa: F401,E402


def train(model, **kwargs):
    """train!(model; kwargs...) -- dispatches on the model type like the reference's methods."""
    if isinstance(model, gpuLDA):
        return train_lda(model, **kwargs)
    if isinstance(model, gpuCTM):
        return train_ctm(model, **kwargs)
    if isinstance(model, gpuCTPF):
        return train_ctpf(model, **kwargs)
    if isinstance(model, gpufLDA):
        return train_flda(model, **kwargs)
    raise TypeError("train!: unsupported model type %r" % type(model).__name__)


def check_model(model):
    """check_model(model) (modelutils.jl:255-360), dispatched on the model type."""
    if isinstance(model, gpuLDA):
        return check_model_lda(model)
    if isinstance(model, gpuCTM):
        return check_model_ctm(model)
    if isinstance(model, gpuCTPF):
        return check_model_ctpf(model)
    if isinstance(model, gpufLDA):
        return check_model_flda(model)
    raise TypeError("check_model: unsupported model type %r" % type(model).__name__)
