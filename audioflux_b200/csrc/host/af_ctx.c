/* af_ctx.c -- device context, error reporting and device-memory helpers (host C over the CUDA
 * runtime API).  There is no CPU compute path in this library: every compute entry point ends
 * in a kernel launch or fails with a recorded message. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "../af_internal.h"

#define AFB200_VERSION 100   /* 0.1.0 */

static __thread char g_err[512];
static __thread int g_device = -1;          /* -1: follow the CUDA current device */
static long long g_launches = 0;

int af_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[audioflux_b200] error %d: %s\n", code, g_err);
    return code;
}

void af_clear_error(void) { g_err[0] = 0; }
const char *afb200_lastError(void) { return g_err; }
int afb200_version(void) { return AFB200_VERSION; }

int af_cuda_check(int e, const char *what) {
    if (e == cudaSuccess) return AF_OK;
    return af_fail(AF_ERR_CUDA, "%s: %s", what, cudaGetErrorString((cudaError_t)e));
}

int afb200_deviceCount(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int afb200_setDevice(int device) {
    int n = afb200_deviceCount();
    if (device < 0 || device >= n) return af_fail(AF_ERR_ARG, "afb200_setDevice(%d): %d device(s) visible", device, n);
    g_device = device;
    return af_cuda_check(cudaSetDevice(device), "cudaSetDevice");
}

int afb200_getDevice(void) {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}

int af_device_ready(void) {
    if (afb200_deviceCount() <= 0)
        return af_fail(AF_ERR_NOGPU, "no CUDA device is visible: libaudioflux_b200 has no CPU fallback");
    if (g_device >= 0) return af_cuda_check(cudaSetDevice(g_device), "cudaSetDevice");
    return AF_OK;
}

int afb200_deviceSynchronize(void) { return af_cuda_check(cudaDeviceSynchronize(), "cudaDeviceSynchronize"); }
long long afb200_kernelLaunchCount(void) { return g_launches; }
void af_count_launch(int n) { __sync_fetch_and_add(&g_launches, (long long)n); }

int af_devbuf_reserve(AfDevBuf *b, size_t bytes) {
    if (b->bytes >= bytes && b->ptr) return AF_OK;
    if (b->ptr) { cudaFree(b->ptr); b->ptr = NULL; b->bytes = 0; }
    if (bytes == 0) return AF_OK;
    int e = cudaMalloc(&b->ptr, bytes);
    if (e != cudaSuccess) { b->ptr = NULL; return af_fail(AF_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString((cudaError_t)e)); }
    b->bytes = bytes;
    return AF_OK;
}

void af_devbuf_free(AfDevBuf *b) { if (b->ptr) cudaFree(b->ptr); b->ptr = NULL; b->bytes = 0; }

/* *dptr must be NULL or an earlier allocation of this call (object fields: calloc'ed); a table uploaded again after a
 * failed lazy initialisation replaces the earlier copy instead of leaking it */
int af_dev_upload(void **dptr, const void *host, size_t bytes) {
    if (*dptr) { cudaFree(*dptr); }
    *dptr = NULL;
    if (bytes == 0) return AF_OK;
    int e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) { *dptr = NULL; return af_fail(AF_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString((cudaError_t)e)); }
    e = cudaMemcpy(*dptr, host, bytes, cudaMemcpyHostToDevice);
    /* A pageable-source cudaMemcpy may return once the bytes sit in the driver's staging buffer, before the DMA has
     * landed; the kernels that read the table run on non-blocking streams, which do not order themselves behind the
     * legacy stream.  Wait for the copy itself (tables are uploaded once per object, never on a hot path). */
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return af_cuda_check(e, "cudaMemcpy H2D (table)");
}

void af_dev_free(void *p) { if (p) cudaFree(p); }

int af_stream_create(void **s) {
    if (*s) return AF_OK;                                  /* already created by an earlier, partially failed initialisation */
    cudaStream_t st;
    int e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { *s = NULL; return af_cuda_check(e, "cudaStreamCreate"); }
    *s = (void *)st;
    return AF_OK;
}
void af_stream_destroy(void *s) { if (s) cudaStreamDestroy((cudaStream_t)s); }
int af_stream_sync(void *s) { return af_cuda_check(cudaStreamSynchronize((cudaStream_t)s), "cudaStreamSynchronize"); }
int af_memcpy_h2d(void *d, const void *h, size_t n, void *s) {
    return af_cuda_check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)s), "cudaMemcpyAsync H2D");
}
int af_memcpy_d2h(void *h, const void *d, size_t n, void *s) {
    return af_cuda_check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)s), "cudaMemcpyAsync D2H");
}
int af_memset_d(void *d, int v, size_t n, void *s) {
    return af_cuda_check(cudaMemsetAsync(d, v, n, (cudaStream_t)s), "cudaMemsetAsync");
}
size_t af_dev_free_bytes(void) {
    size_t f = 0, t = 0;
    if (cudaMemGetInfo(&f, &t) != cudaSuccess) { cudaGetLastError(); return 0; }
    return f;
}
int af_sm_count(void) {
    int d = 0, n = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess) return 0;
    return n;
}

/* ---- events: ordering between the copy / compute / read-back streams of the host-pointer pipelines ---- */
int af_event_create(void **ev) {
    cudaEvent_t e;
    int rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (rc != cudaSuccess) { *ev = NULL; return af_cuda_check(rc, "cudaEventCreate"); }
    *ev = (void *)e;
    return AF_OK;
}
void af_event_destroy(void *ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
int af_event_record(void *ev, void *stream) { return af_cuda_check(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream), "cudaEventRecord"); }
int af_stream_wait_event(void *stream, void *ev) { return af_cuda_check(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0), "cudaStreamWaitEvent"); }

int af_pipe_run(AfPipe *pp, AfChunkFn fn, void *obj, const float *hIn, size_t inPer, int batch,
                float *hOut0, float *hOut1, size_t outPer, void *st) {
    int rc;
    if (!pp->ready) {
        if ((rc = af_stream_create(&pp->inStream)) || (rc = af_stream_create(&pp->outStream))) return rc;
        for (int s = 0; s < 2; s++)
            if ((rc = af_event_create(&pp->evIn[s])) || (rc = af_event_create(&pp->evDone[s])) || (rc = af_event_create(&pp->evOut[s]))) return rc;
        pp->ready = 1;
    }
    /* chunk: about 64 MB of the larger side, a multiple of 16 items when possible, at least 1 */
    const size_t big = inPer > outPer * (hOut1 ? 2 : 1) ? inPer : outPer * (hOut1 ? 2 : 1);
    long long per = ((long long)64 << 20) / (long long)(big * sizeof(float) > 0 ? big * sizeof(float) : 1);
    if (per >= 16) per -= per % 16;
    if (per < 1) per = 1;
    if (per > batch) per = batch;
    const int chunk = (int)per;
    for (int s = 0; s < 2; s++) {
        if ((rc = af_devbuf_reserve(&pp->in[s], sizeof(float) * inPer * chunk)) ||
            (rc = af_devbuf_reserve(&pp->out0[s], sizeof(float) * outPer * chunk))) return rc;
        if (hOut1 && (rc = af_devbuf_reserve(&pp->out1[s], sizeof(float) * outPer * chunk))) return rc;
    }
    int k = 0;
    for (int c0 = 0; c0 < batch; c0 += chunk, k++) {
        const int nb = batch - c0 < chunk ? batch - c0 : chunk, s = k & 1;
        if (k >= 2) {                                   /* slot reuse: its previous transform and read-back are over */
            if ((rc = af_stream_wait_event(pp->inStream, pp->evDone[s])) || (rc = af_stream_wait_event(st, pp->evOut[s]))) return rc;
        }
        if ((rc = af_memcpy_h2d(pp->in[s].ptr, hIn + (size_t)c0 * inPer, sizeof(float) * inPer * nb, pp->inStream))) return rc;
        if ((rc = af_event_record(pp->evIn[s], pp->inStream)) || (rc = af_stream_wait_event(st, pp->evIn[s]))) return rc;
        if ((rc = fn(obj, (const float *)pp->in[s].ptr, nb, (float *)pp->out0[s].ptr, hOut1 ? (float *)pp->out1[s].ptr : NULL, st))) return rc;
        if ((rc = af_event_record(pp->evDone[s], st)) || (rc = af_stream_wait_event(pp->outStream, pp->evDone[s]))) return rc;
        if ((rc = af_memcpy_d2h(hOut0 + (size_t)c0 * outPer, pp->out0[s].ptr, sizeof(float) * outPer * nb, pp->outStream))) return rc;
        if (hOut1 && (rc = af_memcpy_d2h(hOut1 + (size_t)c0 * outPer, pp->out1[s].ptr, sizeof(float) * outPer * nb, pp->outStream))) return rc;
        if ((rc = af_event_record(pp->evOut[s], pp->outStream))) return rc;
    }
    if ((rc = af_stream_sync(pp->inStream)) || (rc = af_stream_sync(st))) return rc;
    return af_stream_sync(pp->outStream);
}

void af_pipe_free(AfPipe *pp) {
    af_stream_destroy(pp->inStream); af_stream_destroy(pp->outStream);
    for (int s = 0; s < 2; s++) {
        af_event_destroy(pp->evIn[s]); af_event_destroy(pp->evDone[s]); af_event_destroy(pp->evOut[s]);
        af_devbuf_free(&pp->in[s]); af_devbuf_free(&pp->out0[s]); af_devbuf_free(&pp->out1[s]);
    }
    memset(pp, 0, sizeof(*pp));
}

/* ---- buffers that other processes (one per GPU) can map: the gathered result of the multi-GPU path ---- */
int afb200_peerAlloc(void **devPtr, size_t bytes) {
    if (!devPtr || bytes == 0) return af_fail(AF_ERR_ARG, "afb200_peerAlloc: bad argument");
    int rc = af_device_ready();
    if (rc) return rc;
    *devPtr = NULL;
    int e = cudaMalloc(devPtr, bytes);          /* plain cudaMalloc: exportable with cudaIpcGetMemHandle at offset 0 */
    if (e != cudaSuccess) { *devPtr = NULL; return af_fail(AF_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString((cudaError_t)e)); }
    return AF_OK;
}
int afb200_peerFree(void *devPtr) { return devPtr ? af_cuda_check(cudaFree(devPtr), "cudaFree") : AF_OK; }
int afb200_ipcGetHandle(void *devPtr, void *handle64) {
    if (!devPtr || !handle64) return af_fail(AF_ERR_ARG, "afb200_ipcGetHandle: bad argument");
    cudaIpcMemHandle_t h;
    int e = cudaIpcGetMemHandle(&h, devPtr);
    if (e != cudaSuccess) return af_cuda_check(e, "cudaIpcGetMemHandle");
    memcpy(handle64, &h, sizeof(h));
    return AF_OK;
}
int afb200_ipcOpenHandle(const void *handle64, void **devPtr) {
    if (!handle64 || !devPtr) return af_fail(AF_ERR_ARG, "afb200_ipcOpenHandle: bad argument");
    int rc = af_device_ready();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    *devPtr = NULL;
    return af_cuda_check(cudaIpcOpenMemHandle(devPtr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
int afb200_ipcCloseHandle(void *devPtr) { return devPtr ? af_cuda_check(cudaIpcCloseMemHandle(devPtr), "cudaIpcCloseMemHandle") : AF_OK; }
