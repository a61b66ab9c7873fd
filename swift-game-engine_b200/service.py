"""CollisionQueryService — the rebuild-vs-refit policy around the query object (host logic only).

Mirror of `final class CollisionQueryService` (Game/SceneServices.swift:33-207): owns the CollisionQuery,
and on every fixed step decides between a full rebuild and a transform refit by diffing a per-entity
snapshot (translation / rotation / scale / vertex count / index count / body type / collides):
  * active set changed, dirty flag, no query yet ............... rebuild   (:52-60)
  * entity count changed, mesh.dirty, snapshot missing, body type or collides changed, vertex or index
    count changed ............................................... rebuild   (:100-163 "structuralChange")
  * squared delta of translation, rotation (quaternion vector) or scale > 1e-6
    ............................................................. updateStaticTransforms / updateDynamicTransforms (:66-75)
Entities are dicts: entity_id, translation (3), rotation (x,y,z,w), scale (3), positions, indices,
body_type (None | "static" | "kinematic" | "dynamic"), collides, dirty, layer, mu_s, mu_k, flatten_ground.
`world_factory(parts)` builds the query object (default: the CUDA CollisionQuery); tests inject a recorder.
"""
import numpy as np

from . import scenes

EPS = np.float32(1e-6)


def _f32(v):
    return np.asarray(v, np.float32)


class CollisionQueryService:
    def __init__(self, world_factory=None):
        if world_factory is None:
            from . import CollisionQuery as world_factory  # the CUDA-backed query (no CPU fallback)
        self._factory = world_factory
        self.query = None
        self._dirty = True
        self._cache = {}
        self._last_active = None
        self.last_action = None  # "rebuild" | "refit" | "none"  (instrumentation, not in the reference)

    def markDirty(self):
        self._dirty = True

    # ---- helpers
    @staticmethod
    def _filter(entities, active_ids):
        out = []
        for e in entities:
            if active_ids is not None and e["entity_id"] not in active_ids:
                continue
            if e.get("collides", True):
                out.append(e)
        return sorted(out, key=lambda e: e["entity_id"])  # deterministic order (the reference: Dictionary order)

    @staticmethod
    def _model(e):
        return scenes.trs_model(e.get("translation", (0, 0, 0)), e.get("rotation", (0, 0, 0, 1)), e.get("scale", (1, 1, 1)))

    @classmethod
    def _part(cls, e):
        bt = e.get("body_type")
        return scenes.part(e["positions"], e["indices"], cls._model(e), layer=e.get("layer", 1), mu_s=e.get("mu_s", 0.8),
                           mu_k=e.get("mu_k", 0.6), flatten_ground=e.get("flatten_ground", False),
                           is_dynamic=(bt is not None and bt != "static"), entity_id=e["entity_id"])

    @staticmethod
    def _snapshot(e):
        return {"translation": _f32(e.get("translation", (0, 0, 0))).copy(), "rotation": _f32(e.get("rotation", (0, 0, 0, 1))).copy(),
                "scale": _f32(e.get("scale", (1, 1, 1))).copy(), "vertex_count": len(np.asarray(e["positions"]).reshape(-1, 3)),
                "index_count": len(np.asarray(e["indices"]).reshape(-1)), "body_type": e.get("body_type"),
                "collides": e.get("collides", True)}

    def _refresh_cache(self, entities, active_ids):  # SceneServices.swift:171-194
        self._cache = {}
        for e in self._filter(entities, active_ids):
            self._cache[e["entity_id"]] = self._snapshot(e)
            e["dirty"] = False

    # ---- rebuild (SceneServices.swift:45-50)
    def rebuild(self, entities, active_ids=None):
        if self.query is not None and hasattr(self.query, "close"):
            self.query.close()
        self.query = self._factory([self._part(e) for e in self._filter(entities, active_ids)])
        self._dirty = False
        self._last_active = None if active_ids is None else set(active_ids)
        self._refresh_cache(entities, active_ids)
        self.last_action = "rebuild"

    # ---- staticMeshChanges (SceneServices.swift:95-169)
    def _changes(self, entities, active_ids):
        ents = self._filter(entities, active_ids)
        if len(ents) != len(self._cache):
            return True, [], []
        static_set, dynamic_set = [], []
        for e in ents:
            if e.get("dirty", False):
                return True, [], []
            snap = self._cache.get(e["entity_id"])
            if snap is None:
                return True, [], []
            bt = e.get("body_type")
            if snap["body_type"] != bt or snap["collides"] != e.get("collides", True):
                return True, [], []
            moved = False
            for key, default in (("translation", (0, 0, 0)), ("rotation", (0, 0, 0, 1)), ("scale", (1, 1, 1))):
                d = _f32(e.get(key, default)) - snap[key]
                if np.float32(np.dot(d, d)) > EPS:
                    moved = True
            if moved:
                (static_set if bt in (None, "static") else dynamic_set).append(e)
            if (len(np.asarray(e["positions"]).reshape(-1, 3)) != snap["vertex_count"]
                    or len(np.asarray(e["indices"]).reshape(-1)) != snap["index_count"]):
                return True, [], []
        return False, static_set, dynamic_set

    # ---- update (SceneServices.swift:52-77)
    def update(self, entities, active_ids=None):
        active = None if active_ids is None else set(active_ids)
        if active != self._last_active or self._dirty or self.query is None:
            self.rebuild(entities, active_ids)
            return
        structural, st, dy = self._changes(entities, active_ids)
        if structural:
            self.rebuild(entities, active_ids)
            return
        self.last_action = "none"
        if st:
            self.query.updateStaticTransforms([e["entity_id"] for e in st], [self._model(e) for e in st])
            self.last_action = "refit"
        if dy:
            self.query.updateDynamicTransforms([e["entity_id"] for e in dy], [self._model(e) for e in dy])
            self.last_action = "refit"
        self._refresh_cache(entities, active_ids)


# ---------------------------------------------------------------- streaming active set (Systems.swift:2354-2411)
CHUNK_SIZE = 512.0  # WorldPosition.chunkSize (Components.swift:55)


def world_to_chunk(world):
    """WorldPosition.fromWorld (Components.swift:58-69): per axis chunk = floor((v + 256) / 512), local = v - chunk*512.
    Returns (chunk int64 (...,3), local float64 (...,3))."""
    w = np.asarray(world, np.float64)
    chunk = np.floor((w + CHUNK_SIZE * 0.5) / CHUNK_SIZE).astype(np.int64)
    return chunk, w - chunk.astype(np.float64) * CHUNK_SIZE


def chunk_to_world(chunk, local):
    """WorldPosition.toWorld (Components.swift:87-93)."""
    return np.asarray(chunk, np.int64).astype(np.float64) * CHUNK_SIZE + np.asarray(local, np.float64)


class ActiveChunkSet:
    """ActiveChunkSystem.fixedUpdate (Systems.swift:2354-2396): the entities whose chunk lies within `radius_chunks`
    (Chebyshev distance, ActiveChunkComponent.radiusChunks default 2, Components.swift:150) of the player's chunk.
    `update` returns (active_entity_ids, active_static_entity_ids); feed the first to CollisionQueryService.update —
    a changed set is what triggers the rebuild there (SceneServices.swift:55-58)."""

    def __init__(self, radius_chunks=2):
        self.radius_chunks = radius_chunks
        self.center_chunk = np.zeros(3, np.int64)
        self.active_entity_ids = set()
        self.active_static_entity_ids = set()

    def update(self, player_world_position, entities):
        """entities: dicts with entity_id and either `chunk` (3 ints) or `world_position` / `translation` (3 floats);
        `body_type` None or "static" with a mesh counts as static (StaticMeshComponent present)."""
        center, _ = world_to_chunk(player_world_position)
        radius = max(int(self.radius_chunks), 0)
        active, active_static = set(), set()
        for e in entities:
            if "chunk" in e:
                chunk = np.asarray(e["chunk"], np.int64)
            else:
                chunk, _ = world_to_chunk(e.get("world_position", e.get("translation", (0, 0, 0))))
            if int(np.abs(chunk - center).max()) <= radius:
                active.add(e["entity_id"])
                if "positions" in e:  # staticStore.contains(e): the entity carries a StaticMeshComponent
                    active_static.add(e["entity_id"])
        self.center_chunk = center
        self.active_entity_ids, self.active_static_entity_ids = active, active_static
        return active, active_static
