"""predict(corp, train_model) -- E-step-only inference on unseen documents with frozen globals
(modelutils.jl:831-855 for LDA/gpuLDA, :886-913 for CTM/gpuCTM): the same device kernels, no M-step."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from .corpus import Corpus, CorpusError, check_corp
from .gpu_ctm import check_model_ctm, gpuCTM, gpufCTM
from .gpu_flda import check_model_flda, gpufLDA
from .gpu_lda import check_model as check_model_lda
from .gpu_lda import gpuLDA


def predict(corp: Corpus, train_model, iter: int = 10, tol: Optional[float] = None, niter: int = 1000, ntol: Optional[float] = None):
    """Returns a new model over ``corp`` whose per-document parameters (gamma / Elogtheta, or lambda / vsq / logzeta)
    were fitted with ``train_model``'s alpha, beta (mu, sigma).  ``iter`` >= 1 sweeps per document."""
    check_corp(corp)
    K = train_model.K
    tol = 1.0 / K**2 if tol is None else tol
    ntol = 1.0 / K**2 if ntol is None else ntol
    if corp.V != train_model.V:
        raise CorpusError("predict corpus and train_model corpus must have identical vocabularies.")
    if tol < 0 or ntol < 0:
        raise ValueError("tolerance parameter must be nonnegative.")
    if iter < 0 or niter < 0:
        raise ValueError("iteration parameter must be nonnegative.")
    if isinstance(train_model, gpuLDA):
        check_model_lda(train_model)
        model = gpuLDA(corp, K)
        model.alpha = np.array(train_model.alpha, dtype=np.float32)
        model.beta = np.array(train_model.beta, dtype=np.float32, order="F")
        model.topics = train_model.topics
        if iter > 0 and corp.flat().nnz > 0:
            model.update_buffer()
            # update_phi!/update_gamma!/update_Elogtheta! per document; nothing is scattered (there is no M-step to feed)
            _lib.check(_lib.load().tmvb_lda_predict(model._handle(), int(iter), float(tol)))
            model.update_host()
        return model
    if isinstance(train_model, gpufLDA):
        # predict(corp, train_model::fLDA) (modelutils.jl:857-884).  The reference copies alpha, beta and topics only -- the new model
        # keeps a freshly drawn kappa and eta = 0.5 -- and its loop reads an undefined `vtol` (:877); here the trained kappa and eta
        # travel with beta (the filter they describe belongs to the trained topics) and `tol` is the stopping tolerance.
        check_model_flda(train_model)
        model = gpufLDA(corp, K)
        model.alpha = np.array(train_model.alpha, dtype=np.float32)
        model.beta = np.array(train_model.beta, dtype=np.float32, order="F")
        model.kappa = np.array(train_model.kappa, dtype=np.float32)
        model.eta = float(train_model.eta)
        model.topics = train_model.topics
        if iter > 0 and corp.flat().nnz > 0:
            model.update_buffer()
            _lib.check(_lib.load().tmvb_flda_predict(model._handle(), int(iter), float(tol)))
            model.update_host()
        return model
    if isinstance(train_model, gpuCTM):
        check_model_ctm(train_model)
        model = gpufCTM(corp, K) if isinstance(train_model, gpufCTM) else gpuCTM(corp, K)
        if isinstance(train_model, gpufCTM):   # predict(corp, train_model::fCTM) (modelutils.jl:915-944), kappa / eta as for fLDA above
            model.kappa = np.array(train_model.kappa, dtype=np.float32)
            model.eta = float(train_model.eta)
        model.mu = np.array(train_model.mu, dtype=np.float32)
        model.sigma = np.array(train_model.sigma, dtype=np.float32)
        model.invsigma = np.array(train_model.invsigma, dtype=np.float32)
        model.beta = np.array(train_model.beta, dtype=np.float32, order="F")
        model.topics = train_model.topics
        if iter > 0 and corp.flat().nnz > 0:
            model.update_buffer()
            _lib.check(_lib.load().tmvb_ctm_predict(model._handle(), int(niter), float(ntol), int(iter), float(tol)))
            model.update_host()
        return model
    raise TypeError("predict: unsupported model type %r" % type(train_model).__name__)


def topicdist(model, d: int) -> np.ndarray:
    """topicdist(model, d) (modelutils.jl:946-983): gamma-normalised topic proportions for LDA, additive-logistic of
    lambda for CTM, Etheta-normalised for CTPF.  ``d`` is 0-based here."""
    if isinstance(model, (gpuLDA, gpufLDA)):
        g = np.asarray(model.gamma[:, d], dtype=np.float64)
        return g / g.sum()
    if isinstance(model, gpuCTM):
        x = np.asarray(model.lam[:, d], dtype=np.float64)
        x = np.exp(x - x.max())
        return x / x.sum()
    g = np.asarray(model.gimel[:, d], dtype=np.float64) / np.asarray(model.dalet, dtype=np.float64)
    return g / g.sum()
