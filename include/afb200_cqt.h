/* afb200_cqt.h -- constant-Q transform.  Replaces /root/reference/src/cqt_algorithm.h:14-62
 * (src/cqt_algorithm.c); chroma/cqcc/cqhc/deconv are "next" rows and not exported yet. */
#ifndef AFB200_CQT_H
#define AFB200_CQT_H
#include "afb200_types.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OpaqueCQT *CQTObj;

int cqtObj_new(CQTObj *cqtObj, int num, int samplate, float minFre, int *isContinue);  /* cqt_algorithm.c:110-120 */
/* :123-247.  -1 if binPerOctave%12 or num%binPerOctave; -2 for isContinue=1 / beta!=0 (VQT). */
int cqtObj_newWith(CQTObj *cqtObj, int num, int *samplate, float *minFre, int *binPerOctave,
                   float *factor, float *beta, float *thresh, WindowType *windowType, int *slideLength,
                   int *isContinue, SpectralFilterBankNormalType *normalType, int *isScale);
int cqtObj_calTimeLength(CQTObj cqtObj, int dataLength);          /* :266-299 */
int cqtObj_getFFTLength(CQTObj cqtObj);                           /* :333-338 */
float *cqtObj_getFreBandArr(CQTObj cqtObj);                       /* :340-343, borrowed */
void cqtObj_setScale(CQTObj cqtObj, int flag);                    /* :458-461 */
/* :463-478.  mRealArr3/mImageArr3: timeLength x num. */
void cqtObj_cqt(CQTObj cqtObj, float *dataArr, int dataLength, float *mRealArr3, float *mImageArr3);
void cqtObj_free(CQTObj cqtObj);

#ifdef __cplusplus
}
#endif
#endif
