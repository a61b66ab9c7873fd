// Stubs.swift — the few engine types Components.swift names but the collision path never touches (renderer / animation
// side of the ECS).  TEST INFRASTRUCTURE: part of the golden-vector harness only.
import simd

public struct GPUMesh {}      // RenderComponent.mesh (GPUMesh.swift wraps Metal buffers)
public struct Material {}     // RenderComponent.material (Material.swift)
public struct Skeleton {}     // SkeletonComponent.skeleton (Skeleton.swift)
public struct MotionProfile { // MotionProfileComponent.profile (Animation.swift); fields the harness never reads
    public var duration: Float = 1
    public init() {}
}
