package de.labathome.cubature;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/** Panama FFM (JDK >= 22) downcall handles for include/cubature_cuda.h. SOURCE ONLY (no JVM in the image). */
final class Native {
    static final Linker LINKER = Linker.nativeLinker();
    static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            System.getProperty("cubature.cuda.lib", "libcubature_cuda.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }

    /* int cubature_cuda_integrand_builtin(int, int, int, const double*, size_t, cub_integrand_t**) */
    static final MethodHandle INTEGRAND_BUILTIN = h("cubature_cuda_integrand_builtin",
            FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    /* int cubature_cuda_integrand_nvrtc(const char*, const char*, int, int, const double*, size_t, cub_integrand_t**, char*, size_t) */
    static final MethodHandle INTEGRAND_NVRTC = h("cubature_cuda_integrand_nvrtc",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, JAVA_LONG));
    /* void cubature_cuda_integrand_free(cub_integrand_t*) */
    static final MethodHandle INTEGRAND_FREE = h("cubature_cuda_integrand_free", FunctionDescriptor.ofVoid(ADDRESS));
    /* int cubature_cuda_integrate(cub_integrand_t*, int, const double*, const double*, const int*, double, double, int,
     *                             int64_t, const cub_options_t*, double*, cub_stats_t*) */
    static final MethodHandle INTEGRATE = h("cubature_cuda_integrate",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_INT,
                    JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    /* int cubature_cuda_checkpoint_info(const char*, cub_checkpoint_info_t*) */
    static final MethodHandle CHECKPOINT_INFO = h("cubature_cuda_checkpoint_info", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    /* int64_t cubature_cuda_checkpoint_regions(const char*, int64_t, int64_t, double*, double*, double*, double*, int32_t*) */
    static final MethodHandle CHECKPOINT_REGIONS = h("cubature_cuda_checkpoint_regions",
            FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    /* int cubature_cuda_rule_points(int, const double*, const double*, double*) */
    static final MethodHandle RULE_POINTS = h("cubature_cuda_rule_points", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    /* const char* cubature_cuda_strerror(int) */
    static final MethodHandle STRERROR = h("cubature_cuda_strerror", FunctionDescriptor.of(ADDRESS, JAVA_INT));

    /** status -> unchecked exception, like the reference (Cubature.java:119-128, Rule.java:164-167). */
    static void check(int status, String detail) {
        if (status == 0) return;
        String msg;
        try {
            MemorySegment s = (MemorySegment) STRERROR.invokeExact(status);
            msg = s.reinterpret(256).getString(0);
        } catch (Throwable t) {
            msg = "status " + status;
        }
        throw new RuntimeException(detail == null || detail.isEmpty() ? msg : msg + ": " + detail);
    }

    private Native() {
    }
}
