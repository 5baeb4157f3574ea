"""`src/scene` of the reference: Camera (render/camera.rs:6-35) and Scene (scene/scene.rs:7-24)."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class Camera:
    eye: tuple = (0.5, 0.5, -500.5)
    target: tuple = (0.5, 0.5, -498.5)
    up: tuple = (0.0, 1.0, 0.0)
    aspect: float = 1.0
    fovy: float = 45.0  # degrees

    @classmethod
    def quick_camera(cls, aspect: float) -> "Camera":
        """camera.rs:16-29."""
        return cls(aspect=aspect)


@dataclass
class Scene:
    """scene.rs:7-24 without the input controller (interactive windowing is out of scope)."""
    width: int = 1600
    height: int = 900  # DEFAULT_SIZE, lib.rs:23
    camera: Camera = field(default=None)

    def __post_init__(self):
        if self.camera is None:
            self.camera = Camera.quick_camera(self.width / self.height)
