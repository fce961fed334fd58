"""Optional-dependency helpers.

Parity: reference ``src/sub/utils/lightning_core_imports.py`` (``RequirementCache``,
``ModuleAvailableCache``, ``LazyModule``, ``requires`` — copied there from lightning-utilities).
Rewritten on ``importlib.metadata``/``packaging`` (no ``pkg_resources``).
"""
from __future__ import annotations

import functools
import importlib
import importlib.util
from importlib import metadata
from types import ModuleType
from typing import Any, Callable, Optional

__all__ = ["RequirementCache", "ModuleAvailableCache", "LazyModule", "requires", "module_available"]


def module_available(module_path: str) -> bool:
    try:
        return importlib.util.find_spec(module_path) is not None
    except (ModuleNotFoundError, ValueError):
        return False


class RequirementCache:
    """``bool(RequirementCache("torch>=2.0"))`` — lazily evaluated, cached; ``str()`` explains."""

    def __init__(self, requirement: str, module: Optional[str] = None) -> None:
        self.requirement, self.module = requirement, module
        self.available: Optional[bool] = None
        self.message = ""

    def _check(self) -> None:
        if self.available is not None:
            return
        try:
            from packaging.requirements import Requirement

            req = Requirement(self.requirement)
            try:
                version = metadata.version(req.name)
                ok = req.specifier.contains(version, prereleases=True) if str(req.specifier) else True
                self.available = bool(ok)
                self.message = f"Requirement {self.requirement!r} {'met' if ok else f'not met: found {version}'}"
            except metadata.PackageNotFoundError:
                self.available = module_available(self.module or req.name.replace("-", "_"))
                self.message = (f"Module {req.name!r} importable" if self.available else
                                f"{self.requirement!r} is not installed. HINT: try `pip install {self.requirement}`")
        except Exception as e:  # noqa: BLE001
            self.available = module_available(self.module or self.requirement)
            self.message = f"{type(e).__name__}: {e}"

    def __bool__(self) -> bool:
        self._check()
        return bool(self.available)

    def __str__(self) -> str:
        self._check()
        return self.message

    __repr__ = __str__


class ModuleAvailableCache(RequirementCache):
    def __init__(self, module: str) -> None:
        super().__init__(module, module)


class LazyModule(ModuleType):
    """Import the real module on first attribute access."""

    def __init__(self, module_name: str, callback: Optional[Callable[[], None]] = None) -> None:
        super().__init__(module_name)
        self._module: Optional[ModuleType] = None
        self._callback = callback

    def _load(self) -> ModuleType:
        if self._module is None:
            if self._callback is not None:
                self._callback()
            self._module = importlib.import_module(self.__name__)
        return self._module

    def __getattr__(self, item: str) -> Any:
        return getattr(self._load(), item)

    def __dir__(self):
        return dir(self._load())


def requires(*module_path_version: str, raise_exception: bool = True) -> Callable:
    """Decorator: the wrapped callable needs these requirements."""

    def decorator(func: Callable) -> Callable:
        reqs = [RequirementCache(r) for r in module_path_version]

        @functools.wraps(func)
        def wrapper(*args: Any, **kwargs: Any) -> Any:
            missing = [str(r) for r in reqs if not r]
            if missing:
                msg = f"Required dependencies not available: {'; '.join(missing)}"
                if raise_exception:
                    raise ModuleNotFoundError(msg)
                import warnings

                warnings.warn(msg)
            return func(*args, **kwargs)

        return wrapper

    return decorator
