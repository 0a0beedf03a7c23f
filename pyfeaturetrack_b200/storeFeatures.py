"""Feature table / feature history bookkeeping for long sequences.

The reference declares `KLT_FeatureHistory` and `KLT_FeatureTable` as empty classes (klt.py:272-283: only the C struct
fields survive, as comments) and has none of the routines that fill them.  This module supplies them with the meaning
they have in the C library the reference was ported from (storeFeatures.c of KLT 1.3.4): a table holds, for every
feature slot, one (x, y, val) record per frame; a history is one slot's records over the frames.  Pure host-side
bookkeeping -- nothing here touches the GPU.  Parity: no reference behaviour exists to compare with; the tests check
round trips.

    ft = KLTCreateFeatureTable(nFrames, nFeatures)
    KLTStoreFeatureList(fl, ft, frame)        # after KLTSelectGoodFeatures / KLTTrackFeatures / KLTReplaceLostFeatures
    KLTExtractFeatureList(fl, ft, frame)
    fh = KLTCreateFeatureHistory(nFrames); KLTExtractFeatureHistory(fh, ft, feat); KLTStoreFeatureHistory(fh, ft, feat)
"""
from .error import KLTError
from .klt import KLT_Feature, KLT_FeatureHistory, KLT_FeatureTable, kltState


def _blank():
    f = KLT_Feature()
    f.x, f.y, f.val = -1.0, -1.0, kltState.KLT_NOT_FOUND
    return f


def KLTCreateFeatureList(nFeatures):
    """A feature list is a plain Python list of KLT_Feature in the reference (klt.py:266-270)."""
    return [KLT_Feature() for _ in range(int(nFeatures))]


def KLTCreateFeatureHistory(nFrames):
    fh = KLT_FeatureHistory()
    fh.nFrames = int(nFrames)
    fh.feature = [_blank() for _ in range(fh.nFrames)]
    return fh


def KLTCreateFeatureTable(nFrames, nFeatures):
    ft = KLT_FeatureTable()
    ft.nFrames, ft.nFeatures = int(nFrames), int(nFeatures)
    ft.feature = [[_blank() for _ in range(ft.nFrames)] for _ in range(ft.nFeatures)]     # feature[slot][frame]
    return ft


def _check_frame(who, ft, frame):
    if frame < 0 or frame >= ft.nFrames:
        KLTError("({0}) Given frame number {1} is not in range of feature table".format(who, frame))


def _check_feat(who, ft, feat):
    if feat < 0 or feat >= ft.nFeatures:
        KLTError("({0}) Given feature number {1} is not in range of feature table".format(who, feat))


def KLTStoreFeatureList(fl, ft, frame):
    _check_frame("KLTStoreFeatureList", ft, frame)
    if len(fl) != ft.nFeatures:
        KLTError("(KLTStoreFeatureList) Feature list and feature table must have the same number of features")
    for slot, feat in zip(ft.feature, fl):
        rec = slot[frame]
        rec.x, rec.y, rec.val = feat.x, feat.y, feat.val


def KLTExtractFeatureList(fl, ft, frame):
    _check_frame("KLTExtractFeatureList", ft, frame)
    if len(fl) != ft.nFeatures:
        KLTError("(KLTExtractFeatureList) Feature list and feature table must have the same number of features")
    for slot, feat in zip(ft.feature, fl):
        rec = slot[frame]
        feat.x, feat.y, feat.val = rec.x, rec.y, rec.val


def KLTStoreFeatureHistory(fh, ft, feat):
    _check_feat("KLTStoreFeatureHistory", ft, feat)
    if fh.nFrames != ft.nFrames:
        KLTError("(KLTStoreFeatureHistory) Feature history and feature table must have the same number of frames")
    for rec, src in zip(ft.feature[feat], fh.feature):
        rec.x, rec.y, rec.val = src.x, src.y, src.val


def KLTExtractFeatureHistory(fh, ft, feat):
    _check_feat("KLTExtractFeatureHistory", ft, feat)
    if fh.nFrames != ft.nFrames:
        KLTError("(KLTExtractFeatureHistory) Feature history and feature table must have the same number of frames")
    for rec, dst in zip(ft.feature[feat], fh.feature):
        dst.x, dst.y, dst.val = rec.x, rec.y, rec.val


def table_arrays(ft):
    """(x[nFeatures, nFrames], y, val) NumPy views of a table: handy for writing tracks out or plotting them."""
    import numpy as np
    x = np.array([[r.x for r in slot] for slot in ft.feature], np.float64).reshape(ft.nFeatures, ft.nFrames)
    y = np.array([[r.y for r in slot] for slot in ft.feature], np.float64).reshape(ft.nFeatures, ft.nFrames)
    v = np.array([[r.val for r in slot] for slot in ft.feature], np.int64).reshape(ft.nFeatures, ft.nFrames)
    return x, y, v
