"""Host mirrors of the reference's ``libs/dataset/normalization/operations.py``
(``normalize_1d`` :21-48 non-``individual`` branch, ``unnormalize_1d`` :50-52).
On the native path these are fused into the lifter kernels
(``egn_lifter_forward``); the numpy forms remain for callers that prepare or
inspect statistics on the host."""


def normalize_1d(data, mean, std, individual=False):
    if individual:
        raise NotImplementedError('individual normalisation is a training-time option')
    return (data - mean) / std


def unnormalize_1d(normalized_data, mean, std):
    return normalized_data * std + mean
