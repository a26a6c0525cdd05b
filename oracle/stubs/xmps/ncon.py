import numpy as np


def ncon(tensors, labels, check_indices=False, **kwargs):
    """ncon semantics: positive labels are contracted, negative labels are the
    output legs ordered -1, -2, ... (gaps allowed)."""
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    names = {}
    for lab in labels:
        for x in lab:
            if x not in names:
                names[x] = letters[len(names)]
    neg = sorted({x for lab in labels for x in lab if x < 0}, reverse=True)
    expr = ",".join("".join(names[x] for x in lab) for lab in labels)
    expr += "->" + "".join(names[x] for x in neg)
    return np.einsum(expr, *tensors)
