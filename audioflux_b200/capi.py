"""ctypes signatures of the audioFlux C API for the time-frequency hot path.

The same table binds ``libaudioflux_b200.so`` (this repo) and any other library
exporting the reference's symbols (e.g. the reference build used as test oracle),
because the drop-in boundary IS this C ABI:

* reference symbols  -- src/stft_algorithm.h:14-40, src/bft_algorithm.h:14-57,
  src/feature/xxcc_algorithm.h:12-39, src/cqt_algorithm.h:14-62,
  src/cwt_algorithm.h:14-45, src/spectrogram_algorithm.h:40-119 (signatures reproduced in include/*.h)
* additive ``*Batch`` / ``afb200_*`` symbols -- include/afb200_ext.h (only bound when present)
"""
from __future__ import annotations

import ctypes as C

P = C.POINTER
c_int_p = P(C.c_int)
c_float_p = P(C.c_float)
vp = C.c_void_p

# name -> (restype, argtypes)
REFERENCE_API = {
    # ---- STFT
    "stftObj_new": (C.c_int, [P(vp), C.c_int, c_int_p, c_int_p, c_int_p]),
    "stftObj_setSlideLength": (None, [vp, C.c_int]),
    "stftObj_enablePadding": (None, [vp, C.c_int]),
    "stftObj_enableContinue": (None, [vp, C.c_int]),
    "stftObj_setPadding": (None, [vp, c_int_p, c_int_p, c_float_p, c_float_p]),
    "stftObj_useWindowDataArr": (None, [vp, vp]),
    "stftObj_getWindowDataArr": (vp, [vp]),
    "stftObj_calTimeLength": (C.c_int, [vp, C.c_int]),
    "stftObj_calDataLength": (C.c_int, [vp, C.c_int]),
    "stftObj_stft": (None, [vp, vp, C.c_int, vp, vp]),
    "stftObj_istft": (None, [vp, vp, vp, C.c_int, C.c_int, vp]),
    "stftObj_free": (None, [vp]),
    # ---- BFT
    "bftObj_new": (C.c_int, [P(vp), C.c_int, C.c_int, c_int_p, c_float_p, c_float_p, c_int_p,
                             c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "bftObj_calTimeLength": (C.c_int, [vp, C.c_int]),
    "bftObj_getFreBandArr": (vp, [vp]),
    "bftObj_getBinBandArr": (vp, [vp]),
    "bftObj_setResultType": (None, [vp, C.c_int]),
    "bftObj_setDataNormValue": (None, [vp, C.c_float]),
    "bftObj_bft": (None, [vp, vp, C.c_int, vp, vp]),
    "bftObj_getTemporalData": (None, [vp, P(c_float_p), P(c_float_p), P(c_float_p)]),
    "bftObj_free": (None, [vp]),
    # ---- XXCC
    "xxccObj_new": (C.c_int, [P(vp), C.c_int]),
    "xxccObj_setTimeLength": (None, [vp, C.c_int]),
    "xxccObj_xxcc": (None, [vp, vp, C.c_int, c_int_p, vp]),
    "xxccObj_xxccStandard": (None, [vp, vp, C.c_int, vp, c_int_p, c_int_p, c_int_p, vp, vp, vp]),
    "xxccObj_free": (None, [vp]),
    # ---- CQT
    "cqtObj_new": (C.c_int, [P(vp), C.c_int, C.c_int, C.c_float, c_int_p]),
    "cqtObj_newWith": (C.c_int, [P(vp), C.c_int, c_int_p, c_float_p, c_int_p, c_float_p, c_float_p,
                                 c_float_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "cqtObj_calTimeLength": (C.c_int, [vp, C.c_int]),
    "cqtObj_getFFTLength": (C.c_int, [vp]),
    "cqtObj_getFreBandArr": (vp, [vp]),
    "cqtObj_setScale": (None, [vp, C.c_int]),
    "cqtObj_cqt": (None, [vp, vp, C.c_int, vp, vp]),
    "cqtObj_chroma": (None, [vp, c_int_p, c_int_p, c_int_p, vp, vp, vp]),
    "cqtObj_cqcc": (None, [vp, vp, C.c_int, c_int_p, vp]),
    "cqtObj_cqhc": (None, [vp, vp, C.c_int, vp]),
    "cqtObj_deconv": (None, [vp, vp, vp, vp]),
    "cqtObj_free": (None, [vp]),
    # ---- CWT
    "cwtObj_new": (C.c_int, [P(vp), C.c_int, C.c_int, c_int_p, c_float_p, c_float_p, c_int_p,
                             c_int_p, c_int_p, c_float_p, c_float_p, c_int_p]),
    "cwtObj_getFreBandArr": (vp, [vp]),
    "cwtObj_getBinBandArr": (vp, [vp]),
    "cwtObj_cwt": (None, [vp, vp, vp, vp]),
    "cwtObj_enableDet": (None, [vp, C.c_int]),
    "cwtObj_cwtDet": (None, [vp, vp, vp, vp]),
    "cwtObj_free": (None, [vp]),
    # ---- PWT (src/pwt_algorithm.h:16-31)
    "pwtObj_new": (C.c_int, [P(vp), C.c_int, C.c_int, c_int_p, c_float_p, c_float_p, c_int_p, c_int_p, c_int_p,
                             c_int_p, c_int_p]),
    "pwtObj_getFreBandArr": (vp, [vp]),
    "pwtObj_getBinBandArr": (vp, [vp]),
    "pwtObj_pwt": (None, [vp, vp, vp, vp]),
    "pwtObj_enableDet": (None, [vp, C.c_int]),
    "pwtObj_pwtDet": (None, [vp, vp, vp, vp]),
    "pwtObj_free": (None, [vp]),
    # reassignment (src/reassign_algorithm.h)
    "reassignObj_new": (C.c_int, [P(vp), C.c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_float_p, c_int_p, c_int_p]),
    "reassignObj_calTimeLength": (C.c_int, [vp, C.c_int]),
    "reassignObj_setResultType": (None, [vp, C.c_int]),
    "reassignObj_setOrder": (None, [vp, C.c_int]),
    "reassignObj_reassign": (None, [vp, vp, C.c_int, vp, vp, vp, vp]),
    "reassignObj_free": (None, [vp]),
    # synchrosqueezing (src/wsst_algorithm.h, src/synsq_algorithm.h)
    "wsstObj_new": (C.c_int, [P(vp), C.c_int, C.c_int, c_int_p, c_float_p, c_float_p, c_int_p, c_int_p, c_int_p,
                              c_float_p, c_float_p, c_float_p, c_int_p]),
    "wsstObj_getFreBandArr": (vp, [vp]),
    "wsstObj_getBinBandArr": (vp, [vp]),
    "wsstObj_setOrder": (None, [vp, C.c_int]),
    "wsstObj_wsst": (None, [vp, vp, vp, vp, vp, vp]),
    "wsstObj_free": (None, [vp]),
    "synsqObj_new": (C.c_int, [P(vp), C.c_int, C.c_int, c_int_p, c_int_p, c_float_p]),
    "synsqObj_synsq": (None, [vp, vp, C.c_int, vp, vp, vp, vp]),
    "synsqObj_free": (None, [vp]),
    # ---- Spectrogram (front door; src/spectrogram_algorithm.h:40-119)
    "spectrogramObj_new": (C.c_int, [P(vp), C.c_int, c_int_p, c_float_p, c_float_p, c_int_p, c_int_p, c_int_p,
                                     c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "spectrogramObj_newLinear": (C.c_int, [P(vp), C.c_int, C.c_int, c_int_p]),
    "spectrogramObj_newMel": (C.c_int, [P(vp), C.c_int, C.c_int, C.c_int, c_int_p]),
    "spectrogramObj_newBark": (C.c_int, [P(vp), C.c_int, C.c_int, C.c_int, c_int_p]),
    "spectrogramObj_newErb": (C.c_int, [P(vp), C.c_int, C.c_int, C.c_int, c_int_p]),
    "spectrogramObj_newChroma": (C.c_int, [P(vp), C.c_int, C.c_int, c_int_p]),
    "spectrogramObj_newDeep": (C.c_int, [P(vp), C.c_int, C.c_int, C.c_int, c_int_p]),
    "spectrogramObj_newDeepChroma": (C.c_int, [P(vp), C.c_int, C.c_int, c_int_p]),
    "spectrogramObj_enableDebug": (None, [vp, C.c_int]),
    "spectrogramObj_setDataNormValue": (None, [vp, C.c_float]),
    "spectrogramObj_calTimeLength": (C.c_int, [vp, C.c_int]),
    "spectrogramObj_getFreBandArr": (vp, [vp]),
    "spectrogramObj_getBinBandArr": (vp, [vp]),
    "spectrogramObj_getBandNum": (C.c_int, [vp]),
    "spectrogramObj_getBinBandLength": (C.c_int, [vp]),
    "spectrogramObj_spectrogram": (None, [vp, vp, C.c_int, vp, vp]),
    "spectrogramObj_xxcc": (None, [vp, vp, C.c_int, c_int_p, vp]),
    "spectrogramObj_mfcc": (None, [vp, vp, C.c_int, vp]),
    "spectrogramObj_deconv": (None, [vp, vp, vp, vp]),
    "spectrogramObj_bfcc": (None, [vp, vp, C.c_int, vp]),
    "spectrogramObj_gtcc": (None, [vp, vp, C.c_int, vp]),
    "spectrogramObj_lfcc": (None, [vp, vp, C.c_int, vp]),
    "spectrogramObj_free": (None, [vp]),
}

# additive entry points of libaudioflux_b200.so (include/afb200_ext.h)
EXTENSION_API = {
    "afb200_version": (C.c_int, []),
    "afb200_deviceCount": (C.c_int, []),
    "afb200_setDevice": (C.c_int, [C.c_int]),
    "afb200_getDevice": (C.c_int, []),
    "afb200_lastError": (C.c_char_p, []),
    "afb200_kernelLaunchCount": (C.c_longlong, []),
    "afb200_deviceSynchronize": (C.c_int, []),
    "stftObj_stftBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]),
    "stftObj_istftBatch": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]),
    "bftObj_bftBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]),
    "bftObj_mfccBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]),
    "bftObj_getFilterBankArr": (C.c_int, [vp, vp]),
    "bftObj_mfccPlanMode": (C.c_int, [vp]),
    "bftObj_mfccBatchScatter": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, P(vp), vp]),
    "afb200_peerAlloc": (C.c_int, [P(vp), C.c_size_t]),
    "afb200_peerFree": (C.c_int, [vp]),
    "afb200_ipcGetHandle": (C.c_int, [vp, vp]),
    "afb200_ipcOpenHandle": (C.c_int, [vp, P(vp)]),
    "afb200_ipcCloseHandle": (C.c_int, [vp]),
    "xxccObj_xxccBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]),
    "cqtObj_cqtBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]),
    "cqtObj_getKernelBank": (C.c_int, [vp, vp, vp]),
    "cqtObj_chromaBatch": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]),
    "cqtObj_cqccBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]),
    "cqtObj_cqhcBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp]),
    "cqtObj_deconvBatch": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "xxccObj_xxccStandardBatch": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            vp, vp, vp, C.c_int, vp]),
    "spectrogramObj_spectrogramBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]),
    "spectrogramObj_mfccBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]),
    "spectrogramObj_deconvBatch": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "cwtObj_cwtBatch": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "cwtObj_cwtDetBatch": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "cwtObj_getFilterBankArr": (C.c_int, [vp, vp]),
    "pwtObj_pwtBatch": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "pwtObj_pwtDetBatch": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "pwtObj_getFilterBankArr": (C.c_int, [vp, vp]),
    "wsstObj_wsstDevice": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "reassignObj_reassignBatch": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp]),
    "synsqObj_synsqDevice": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, vp]),
    "afb200_window": (C.c_int, [C.c_int, C.c_int, vp]),
    "afb200_auditoryFilterBank": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_float, C.c_float, C.c_int, vp, vp, vp]),
    "afb200_decimatorTaps": (C.c_int, [vp, vp]),
    "afb200_mfccIntervalPlan": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]),
    "afb200_mfccBankPlan2": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "afb200_chromaCqtFilterBank": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, vp]),
}

# setup-time builders exported (non-static) by the reference only; used by tests to
# compare constant tables (src/dsp/flux_window.h, src/filterbank/*.h)
REFERENCE_BUILDERS = {
    "window_calFFTWindow": (vp, [C.c_int, C.c_int]),
    "auditory_filterBank": (None, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.c_int, vp, vp, vp]),
    "chroma_cqtFilterBank": (None, [C.c_int, C.c_int, C.c_int, c_float_p, vp]),
}


def bind(lib: C.CDLL, tables=(REFERENCE_API, EXTENSION_API, REFERENCE_BUILDERS)) -> dict:
    """Apply argtypes/restype for every symbol the library actually exports.
    Returns {name: bool present}."""
    present = {}
    for table in tables:
        for name, (res, args) in table.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                present[name] = False
                continue
            fn.restype = res
            fn.argtypes = args
            present[name] = True
    return present


def opt_int(v):
    return None if v is None else C.byref(C.c_int(int(v)))


def opt_float(v):
    return None if v is None else C.byref(C.c_float(float(v)))
