package de.labathome.cubature;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

/**
 * A device-side integrand plugin: the GPU replacement of the reference's
 * {@code UnaryOperator<double[][]>} (reference Cubature.java:98-99). Either one of the built-in families
 * of cubature_b200/csrc/integrands.h or CUDA C source compiled for sm_100a with NVRTC.
 *
 * SOURCE ONLY: there is no JVM in the build image, so this file has never been compiled; it is kept
 * deliberately thin (every call is one downcall into libcubature_cuda.so) and is mirrored line for
 * line by the Python binding cubature_b200/__init__.py, which the test-suite exercises.
 */
public final class DeviceIntegrand implements AutoCloseable {
    public static final int GENZ_OSCILLATORY = 1, GENZ_PRODUCT_PEAK = 2, GENZ_CORNER_PEAK = 3, GENZ_GAUSSIAN = 4,
            GENZ_C0 = 5, GENZ_DISCONTINUOUS = 6, GAUSS_ISO = 10, REFTEST = 20, OSC_VEC = 30, CEXP = 31;

    final MemorySegment handle;
    final int dim, fdim;

    private DeviceIntegrand(MemorySegment handle, int dim, int fdim) {
        this.handle = handle;
        this.dim = dim;
        this.fdim = fdim;
    }

    /** cubature_cuda_integrand_builtin */
    public static DeviceIntegrand builtin(int family, int dim, int fdim, double... params) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment p = params.length == 0 ? MemorySegment.NULL : a.allocateFrom(ValueLayout.JAVA_DOUBLE, params);
            MemorySegment out = a.allocate(ValueLayout.ADDRESS);
            int st = (int) Native.INTEGRAND_BUILTIN.invokeExact(family, dim, fdim, p, (long) params.length, out);
            Native.check(st, "");
            return new DeviceIntegrand(out.get(ValueLayout.ADDRESS, 0), dim, fdim);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /**
     * cubature_cuda_integrand_nvrtc: {@code cudaSource} must define
     * {@code __device__ void <entry>(const double* x, double* f, const double* p)} with x[CUB_DIM], f[CUB_FDIM].
     */
    public static DeviceIntegrand fromSource(String cudaSource, String entry, int dim, int fdim, double... params) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment src = a.allocateFrom(cudaSource);
            MemorySegment ent = a.allocateFrom(entry);
            MemorySegment p = params.length == 0 ? MemorySegment.NULL : a.allocateFrom(ValueLayout.JAVA_DOUBLE, params);
            MemorySegment out = a.allocate(ValueLayout.ADDRESS);
            MemorySegment log = a.allocate(16384);
            int st = (int) Native.INTEGRAND_NVRTC.invokeExact(src, ent, dim, fdim, p, (long) params.length, out, log, 16384L);
            Native.check(st, log.getString(0));
            return new DeviceIntegrand(out.get(ValueLayout.ADDRESS, 0), dim, fdim);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    @Override
    public void close() {
        try {
            Native.INTEGRAND_FREE.invokeExact(handle);
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }
}
