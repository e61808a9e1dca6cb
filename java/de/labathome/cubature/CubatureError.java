package de.labathome.cubature;

/**
 * Error norms for vector-valued integrands. Same constants in the same order as the reference's enum
 * (reference: src/main/java/de/labathome/cubature/CubatureError.java:9-29); the ordinal is what crosses
 * the C ABI (CUB_NORM_* in include/cubature_cuda.h).
 */
public enum CubatureError {
    INDIVIDUAL, PAIRED, L2, L1, LINF
}
