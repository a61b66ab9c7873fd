// main.swift — golden-vector harness: runs the UNMODIFIED reference sources (Game/CollisionQuery.swift, the physics part of
// Game/Systems.swift, Game/Components.swift, Game/World.swift ...) on the inputs export_inputs.py wrote and dumps the
// results in the record layouts of include/cq.h, for import_goldens.py to turn into tests/golden/swift_*.npz.
// TEST INFRASTRUCTURE (oracle/): never part of the product.  Written without a Swift toolchain at hand (none exists in
// the build image): build.sh documents what to do if a name has drifted.
//
//   usage: cq_swift_ref <inputs.bin> <outputs.bin>
//
// Input (little endian): "CQSW" u32 version | u32 nParts | parts | u32 nScenarios | scenarios.
//   part:     u32 entityTag, u32 layer, f32 muS, f32 muK, u8 flatten, u8 dynamic, u16 0, u32 nVerts, u32 nIdx,
//             nVerts x 3 f32 WORLD-space positions (the entity gets an identity transform: simd_mul(identity, p) == p
//             exactly, so the triangles are bit-identical to the ones the library and the C++ oracle consume), nIdx u32
//   scenario: u32 kind, then
//     1 casts      u32 mode (0 all, 1 blocking, 2 ground), u32 n, n x {from f32x3, delta f32x3, radius, halfHeight, mask u32, minNormalY}
//     2 overlap    u32 n, n x {from f32x3, radius, halfHeight, mask u32}
//     3 overlapAll u32 maxHits, u32 n, n x capsule
//     4 rays       u32 n, n x {origin f32x3, direction f32x3, maxDistance, mask u32}
//     5 walk       u32 n, u32 steps, f32 dt, f32 gravity x3, u32 applyGravity, f32 maxAccel, f32 maxDecel, u32 hasIntent,
//                  13 x 4 B controller parameters (cq_controller_params), n x {position f32x3, velocity f32x3},
//                  if hasIntent: steps x n x desiredVelocity f32x3
// Output: "CQSO" u32 version | u32 nStatic, nStatic x u32 entityTag (the order TriangleMeshSet.rebuild saw the static
//   entities in: Swift Dictionary order — the library is given its parts in THAT order so that triangle numbering
//   agrees) | u32 nDynamic, ... | per scenario its records: cast hits / overlap hits (44 B), overlapAll: n x maxHits
//   hits + n x i32 counts, ray hits (32 B), walk: steps x n x cq_character_state (168 B).
import Foundation
import simd

struct Reader {
    let d: Data
    var p = 0
    mutating func u32() -> UInt32 { defer { p += 4 }; return d.subdata(in: p..<p + 4).withUnsafeBytes { $0.loadUnaligned(as: UInt32.self) } }
    mutating func f32() -> Float { Float(bitPattern: u32()) }
    mutating func u8() -> UInt8 { defer { p += 1 }; return d[p] }
    mutating func v3() -> SIMD3<Float> { SIMD3<Float>(f32(), f32(), f32()) }
}
struct Writer {
    var d = Data()
    mutating func u32(_ v: UInt32) { var x = v.littleEndian; withUnsafeBytes(of: &x) { d.append(contentsOf: $0) } }
    mutating func i32(_ v: Int32) { u32(UInt32(bitPattern: v)) }
    mutating func f32(_ v: Float) { u32(v.bitPattern) }
    mutating func f64(_ v: Double) { var x = v.bitPattern.littleEndian; withUnsafeBytes(of: &x) { d.append(contentsOf: $0) } }
    mutating func u8(_ v: UInt8) { d.append(v) }
    mutating func v3(_ v: SIMD3<Float>) { f32(v.x); f32(v.y); f32(v.z) }
}

let args = CommandLine.arguments
guard args.count == 3, let blob = FileManager.default.contents(atPath: args[1]) else {
    FileHandle.standardError.write("usage: cq_swift_ref <inputs.bin> <outputs.bin>\n".data(using: .utf8)!)
    exit(2)
}
var r = Reader(d: blob)
precondition(r.u32() == 0x5753_5143 /* "CQSW" */ && r.u32() == 1, "bad input file")

// ---- world: one entity per part, identity transform, world-space vertices
let world = World()
var tagOf: [Entity: UInt32] = [:]
let nParts = Int(r.u32())
for _ in 0..<nParts {
    let tag = r.u32(), layer = r.u32()
    let muS = r.f32(), muK = r.f32()
    let flatten = r.u8() != 0, dynamic = r.u8() != 0
    _ = r.u8(); _ = r.u8()
    let nVerts = Int(r.u32()), nIdx = Int(r.u32())
    var pos: [SIMD3<Float>] = []
    pos.reserveCapacity(nVerts)
    for _ in 0..<nVerts { pos.append(r.v3()) }
    var idx: [UInt32] = []
    idx.reserveCapacity(nIdx)
    for _ in 0..<nIdx { idx.append(r.u32()) }
    let mesh = ProceduralMeshDescriptor(streams: VertexStreams(positions: pos), indices32: idx)
    let e = world.createEntity()
    tagOf[e] = tag
    world.add(e, TransformComponent())
    world.add(e, StaticMeshComponent(mesh: mesh, material: SurfaceMaterial(muS: muS, muK: muK, flattenGround: flatten),
                                     collisionLayer: layer))
    if dynamic { world.add(e, PhysicsBodyComponent(bodyType: .kinematic)) }
}
let query = CollisionQuery(world: world)

var w = Writer()
w.u32(0x4F53_5143) // "CQSO"
w.u32(1)
do { // the entity order the reference's rebuild saw (same dictionary, not mutated since): statics, then dynamics
    let seen = world.query(TransformComponent.self, StaticMeshComponent.self)
    let bodies = world.store(PhysicsBodyComponent.self)
    let statics = seen.filter { bodies[$0] == nil || bodies[$0]!.bodyType == .static }
    let dynamics = seen.filter { !(bodies[$0] == nil || bodies[$0]!.bodyType == .static) }
    for group in [statics, dynamics] {
        w.u32(UInt32(group.count))
        for e in group { w.u32(tagOf[e]!) }
    }
}

func put(_ h: CapsuleCastHit?) {
    if let h { w.f32(h.toi); w.v3(h.position); w.v3(h.normal); w.v3(h.triangleNormal); w.i32(Int32(h.triangleIndex)) }
    else { for _ in 0..<10 { w.f32(0) }; w.i32(-1) }
}
func put(_ h: CapsuleOverlapHit?) {
    if let h { w.f32(h.depth); w.v3(h.position); w.v3(h.normal); w.v3(h.triangleNormal); w.i32(Int32(h.triangleIndex)) }
    else { for _ in 0..<10 { w.f32(0) }; w.i32(-1) }
}
func put(_ h: RaycastHit?) {
    if let h { w.f32(h.distance); w.v3(h.position); w.v3(h.normal); w.i32(Int32(h.triangleIndex)) }
    else { for _ in 0..<7 { w.f32(0) }; w.i32(-1) }
}

let nScenarios = Int(r.u32())
for _ in 0..<nScenarios {
    switch r.u32() {
    case 1:
        let mode = r.u32(), n = Int(r.u32())
        for _ in 0..<n {
            let from = r.v3(), delta = r.v3(), radius = r.f32(), hh = r.f32(), mask = r.u32(), minY = r.f32()
            switch mode {
            case 0: put(query.capsuleCast(from: from, delta: delta, radius: radius, halfHeight: hh, mask: mask))
            case 1: put(query.capsuleCastBlocking(from: from, delta: delta, radius: radius, halfHeight: hh, mask: mask))
            default: put(query.capsuleCastGround(from: from, delta: delta, radius: radius, halfHeight: hh, minNormalY: minY, mask: mask))
            }
        }
    case 2:
        let n = Int(r.u32())
        for _ in 0..<n {
            let from = r.v3(), radius = r.f32(), hh = r.f32(), mask = r.u32()
            put(query.capsuleOverlap(from: from, radius: radius, halfHeight: hh, mask: mask))
        }
    case 3:
        let maxHits = Int(r.u32()), n = Int(r.u32())
        var counts: [Int32] = []
        for _ in 0..<n {
            let from = r.v3(), radius = r.f32(), hh = r.f32(), mask = r.u32()
            let hits = query.capsuleOverlapAll(from: from, radius: radius, halfHeight: hh, maxHits: maxHits, mask: mask)
            counts.append(Int32(hits.count))
            for k in 0..<max(1, maxHits) { put(k < hits.count ? hits[k] : nil as CapsuleOverlapHit?) }
        }
        for c in counts { w.i32(c) }
    case 4:
        let n = Int(r.u32())
        for _ in 0..<n {
            let o = r.v3(), d = r.v3(), maxD = r.f32(), mask = r.u32()
            put(query.raycast(origin: o, direction: d, maxDistance: maxD, mask: mask))
        }
    case 5:
        let n = Int(r.u32()), steps = Int(r.u32()), dt = r.f32()
        let gravity = r.v3(), applyGravity = r.u32() != 0
        let maxAccel = r.f32(), maxDecel = r.f32(), hasIntent = r.u32() != 0
        var c = CharacterControllerComponent()
        c.radius = r.f32(); c.halfHeight = r.f32(); c.skinWidth = r.f32(); c.groundSnapSkin = r.f32(); c.snapDistance = r.f32()
        c.fallProbeDistance = r.f32(); c.groundSnapMaxSpeed = r.f32(); c.groundSnapMaxToi = r.f32(); c.groundSnapMaxStep = r.f32()
        c.groundSweepMaxStep = r.f32(); c.maxSlideIterations = Int(Int32(bitPattern: r.u32())); c.minGroundDot = r.f32()
        c.collisionMask = r.u32()
        var chars: [Entity] = []
        for _ in 0..<n {
            let p = r.v3(), v = r.v3()
            let e = world.createEntity()
            world.add(e, PhysicsBodyComponent(bodyType: .dynamic, position: p, linearVelocity: v))
            world.add(e, c)
            var move = MovementComponent()
            move.maxAcceleration = maxAccel
            move.maxDeceleration = maxDecel
            world.add(e, move)
            chars.append(e)
        }
        let intent = PhysicsIntentSystem(), grav = GravitySystem(gravity: gravity), mover = KinematicMoveStopSystem(gravity: gravity)
        mover.setQuery(query)
        let bodies = world.store(PhysicsBodyComponent.self), ctrls = world.store(CharacterControllerComponent.self)
        for _ in 0..<steps {
            if hasIntent {
                for e in chars { world.add(e, MoveIntentComponent(desiredVelocity: r.v3())) }
                intent.fixedUpdate(world: world, dt: dt)
            }
            if applyGravity { grav.fixedUpdate(world: world, dt: dt) }
            mover.fixedUpdate(world: world, dt: dt)
            for e in chars { // cq_character_state, 168 bytes
                let b = bodies[e]!, k = ctrls[e]!
                w.f64(b.position.x); w.f64(b.position.y); w.f64(b.position.z)
                w.f64(b.linearVelocity.x); w.f64(b.linearVelocity.y); w.f64(b.linearVelocity.z)
                w.v3(k.groundNormal); w.f32(k.groundDistance); w.v3(k.sideContactNormal)
                w.i32(Int32(k.groundTriangleIndex)); w.i32(Int32(k.groundTransitionFrames)); w.i32(Int32(k.sideContactFrames))
                w.i32(Int32(k.contactManifoldFrames)); w.i32(Int32(k.contactManifoldTriangles.count))
                for j in 0..<4 { w.i32(j < k.contactManifoldTriangles.count ? Int32(k.contactManifoldTriangles[j]) : 0) }
                for j in 0..<4 { w.v3(j < k.contactManifoldNormals.count ? k.contactManifoldNormals[j] : .zero) }
                w.u8(k.grounded ? 1 : 0); w.u8(k.groundedNear ? 1 : 0); w.u8(k.groundSliding ? 1 : 0)
                for _ in 0..<5 { w.u8(0) }
            }
        }
        for e in chars { world.destroyEntity(e) } // the next scenario starts from the static world again
    default:
        fatalError("unknown scenario kind")
    }
}
try! w.d.write(to: URL(fileURLWithPath: args[2]))
