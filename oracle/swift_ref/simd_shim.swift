// simd_shim.swift — stand-in for Apple's `simd` module on Linux, compiled as a Swift module NAMED `simd` so that the
// reference's sources (`import simd`) build unmodified.  TEST INFRASTRUCTURE (oracle/): used only by the golden-vector
// harness in this directory; never part of the product.
//
// Only what the hot path touches is here (census: SURVEY.md §8c / BASELINE.md §3): dot, cross, length, length_squared,
// normalize, min, max on SIMD3<Float> / SIMD3<Double>, 4x4 matrices (columns init, quaternion init, mul), simd_quatf
// (angle-axis init).  Definitions are the ones BASELINE.md §3 documents and both sides of the parity check already
// share: left-to-right sums, IEEE sqrt / divide, normalize = x * (1 / sqrt(len2)), component-wise min / max.
// Swift does not contract a * b + c into an FMA, so these are the same IEEE operation sequences as the C++ oracle
// (g++ -ffp-contract=off) and the CUDA kernels (nvcc -fmad=false).  Apple's own implementation is not inspectable here
// and may differ in the last ulp on arm64.

public typealias simd_float3 = SIMD3<Float>
public typealias simd_float4 = SIMD4<Float>
public typealias simd_double3 = SIMD3<Double>

@inlinable public func simd_dot(_ a: SIMD3<Float>, _ b: SIMD3<Float>) -> Float { (a.x * b.x + a.y * b.y) + a.z * b.z }
@inlinable public func simd_dot(_ a: SIMD3<Double>, _ b: SIMD3<Double>) -> Double { (a.x * b.x + a.y * b.y) + a.z * b.z }
@inlinable public func simd_dot(_ a: SIMD4<Float>, _ b: SIMD4<Float>) -> Float { ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w }

@inlinable public func simd_cross(_ a: SIMD3<Float>, _ b: SIMD3<Float>) -> SIMD3<Float> {
    SIMD3<Float>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x)
}
@inlinable public func simd_cross(_ a: SIMD3<Double>, _ b: SIMD3<Double>) -> SIMD3<Double> {
    SIMD3<Double>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x)
}

@inlinable public func simd_length_squared(_ a: SIMD3<Float>) -> Float { simd_dot(a, a) }
@inlinable public func simd_length_squared(_ a: SIMD3<Double>) -> Double { simd_dot(a, a) }
@inlinable public func simd_length(_ a: SIMD3<Float>) -> Float { simd_dot(a, a).squareRoot() }
@inlinable public func simd_length(_ a: SIMD3<Double>) -> Double { simd_dot(a, a).squareRoot() }
@inlinable public func simd_normalize(_ a: SIMD3<Float>) -> SIMD3<Float> { a * (1.0 / simd_dot(a, a).squareRoot()) }
@inlinable public func simd_normalize(_ a: SIMD3<Double>) -> SIMD3<Double> { a * (1.0 / simd_dot(a, a).squareRoot()) }

// fminf / fmaxf per component (no NaNs occur on the path)
@inlinable public func simd_min(_ a: SIMD3<Float>, _ b: SIMD3<Float>) -> SIMD3<Float> {
    SIMD3<Float>(b.x < a.x ? b.x : a.x, b.y < a.y ? b.y : a.y, b.z < a.z ? b.z : a.z)
}
@inlinable public func simd_max(_ a: SIMD3<Float>, _ b: SIMD3<Float>) -> SIMD3<Float> {
    SIMD3<Float>(b.x > a.x ? b.x : a.x, b.y > a.y ? b.y : a.y, b.z > a.z ? b.z : a.z)
}

public struct simd_quatf {
    public var vector: SIMD4<Float> // (ix, iy, iz, r)
    public init(ix: Float, iy: Float, iz: Float, r: Float) { vector = SIMD4<Float>(ix, iy, iz, r) }
    public init(vector: SIMD4<Float>) { self.vector = vector }
    public init(angle: Float, axis: SIMD3<Float>) {
        let h = angle * 0.5
        let s = Float(_sinD(Double(h))), c = Float(_cosD(Double(h)))
        vector = SIMD4<Float>(axis.x * s, axis.y * s, axis.z * s, c)
    }
    public var real: Float { vector.w }
    public var imag: SIMD3<Float> { SIMD3<Float>(vector.x, vector.y, vector.z) }
}
// Hamilton product
public func simd_mul(_ a: simd_quatf, _ b: simd_quatf) -> simd_quatf {
    let av = a.imag, bv = b.imag
    let v = bv * a.real + av * b.real + simd_cross(av, bv)
    return simd_quatf(ix: v.x, iy: v.y, iz: v.z, r: a.real * b.real - simd_dot(av, bv))
}

public struct simd_float4x4 {
    public var columns: (SIMD4<Float>, SIMD4<Float>, SIMD4<Float>, SIMD4<Float>)
    public init(columns: (SIMD4<Float>, SIMD4<Float>, SIMD4<Float>, SIMD4<Float>)) { self.columns = columns }
    public init(_ c0: SIMD4<Float>, _ c1: SIMD4<Float>, _ c2: SIMD4<Float>, _ c3: SIMD4<Float>) { columns = (c0, c1, c2, c3) }
    public init(diagonal d: SIMD4<Float>) {
        columns = (SIMD4<Float>(d.x, 0, 0, 0), SIMD4<Float>(0, d.y, 0, 0), SIMD4<Float>(0, 0, d.z, 0), SIMD4<Float>(0, 0, 0, d.w))
    }
    /// rotation matrix of a unit quaternion (column-major), the standard formula
    public init(_ q: simd_quatf) {
        let x = q.vector.x, y = q.vector.y, z = q.vector.z, w = q.vector.w
        let xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z
        columns = (SIMD4<Float>(1 - 2 * (yy + zz), 2 * (xy + wz), 2 * (xz - wy), 0),
                   SIMD4<Float>(2 * (xy - wz), 1 - 2 * (xx + zz), 2 * (yz + wx), 0),
                   SIMD4<Float>(2 * (xz + wy), 2 * (yz - wx), 1 - 2 * (xx + yy), 0),
                   SIMD4<Float>(0, 0, 0, 1))
    }
}
public typealias matrix_float4x4 = simd_float4x4
public let matrix_identity_float4x4 = simd_float4x4(diagonal: SIMD4<Float>(1, 1, 1, 1))

// ((c0 * v.x + c1 * v.y) + c2 * v.z) + c3 * v.w
@inlinable public func simd_mul(_ m: simd_float4x4, _ v: SIMD4<Float>) -> SIMD4<Float> {
    ((m.columns.0 * v.x + m.columns.1 * v.y) + m.columns.2 * v.z) + m.columns.3 * v.w
}
public func simd_mul(_ a: simd_float4x4, _ b: simd_float4x4) -> simd_float4x4 {
    simd_float4x4(simd_mul(a, b.columns.0), simd_mul(a, b.columns.1), simd_mul(a, b.columns.2), simd_mul(a, b.columns.3))
}

// sin / cos without Foundation (the shim must not pull in Glibc differences): forwarded to the platform libm
#if canImport(Glibc)
import Glibc
@usableFromInline func _sinD(_ x: Double) -> Double { Glibc.sin(x) }
@usableFromInline func _cosD(_ x: Double) -> Double { Glibc.cos(x) }
#else
import Darwin
@usableFromInline func _sinD(_ x: Double) -> Double { Darwin.sin(x) }
@usableFromInline func _cosD(_ x: Double) -> Double { Darwin.cos(x) }
#endif
