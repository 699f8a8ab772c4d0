"""Per-model sample caps shared by SAP and MOSAP (sap.py:222-240, mosap.py:326-344): a cap on the number
of samples of model i is the linear constraint  ES[i] . m <= cap_i  on the group sample vector m."""
import numpy as np


def max_sample_constraints(ES, n_models, max_model_samples):
    """(rows, caps): one indicator row ES[i] and one integer cap per model with a finite entry in
    ``max_model_samples``; two empty lists when no caps are given.  Same argument checks (and messages) as
    the reference: a numpy array of length N whose first entry allows at least one high-fidelity sample."""
    if max_model_samples is None:
        return [], []
    well_formed = isinstance(max_model_samples, np.ndarray) and len(max_model_samples) == n_models
    if not well_formed:
        raise ValueError("The maximum number of model samples must be prescribed as a numpy array of the same length as the number of models.")
    if max_model_samples[0] < 1:
        raise ValueError("The high-fidelity model must be sampled at least once.")
    capped = [i for i in range(n_models) if np.isfinite(max_model_samples[i])]
    rows = [ES[i] for i in capped]
    caps = [int(np.round(max_model_samples[i])) for i in capped]
    return rows, caps
