package de.labathome.cubature;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

/**
 * Drop-in facade with the reference's entry point (reference Cubature.java:108-170):
 * {@code double[][] integrate(integrand, xmin, xmax, relTol, absTol, norm, maxEval)} returning
 * {@code [2][fdim]} = values, error estimates (reference Cubature.java:169). The only difference a caller
 * sees is the type of the first argument: the integrand has to run on the GPU, so it is a
 * {@link DeviceIntegrand} (built-in family or CUDA C source) instead of a host lambda, and
 * {@code maxEval} is a long (BASELINE config 5 needs 1e11; the reference's int cannot express it).
 *
 * SOURCE ONLY: never compiled here (no JVM in the image). Mirrored by cubature_b200.Cubature (Python).
 */
public final class Cubature {

    public static double[][] integrate(DeviceIntegrand integrand, double[] xmin, double[] xmax, double relTol,
            double absTol, CubatureError norm, long maxEval) {
        if (xmin == null || xmin.length == 0) {
            throw new RuntimeException("xmin cannot be null or empty");            // reference Cubature.java:119-123
        }
        if (xmax == null || xmax.length != xmin.length) {
            throw new RuntimeException("xmin and xmax must have the same length"); // reference Cubature.java:125-128
        }
        final int dim = xmin.length, fdim = integrand.fdim;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment lo = a.allocateFrom(ValueLayout.JAVA_DOUBLE, xmin);
            MemorySegment hi = a.allocateFrom(ValueLayout.JAVA_DOUBLE, xmax);
            MemorySegment out = a.allocate(ValueLayout.JAVA_DOUBLE, 2L * fdim);
            // transform = NULL: infinite bounds are mapped automatically; options = NULL: all-device mode, current GPU;
            // stats = NULL. See include/cubature_cuda.h for cub_options_t / cub_stats_t when they are wanted.
            int st = (int) Native.INTEGRATE.invokeExact(integrand.handle, dim, lo, hi, MemorySegment.NULL, relTol, absTol,
                    norm.ordinal(), maxEval, MemorySegment.NULL, out, MemorySegment.NULL);
            Native.check(st, "");
            double[] flat = out.toArray(ValueLayout.JAVA_DOUBLE);
            double[] val = new double[fdim], err = new double[fdim];
            System.arraycopy(flat, 0, val, 0, fdim);
            System.arraycopy(flat, fdim, err, 0, fdim);
            return new double[][] { val, err };
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    private Cubature() {
    }
}
