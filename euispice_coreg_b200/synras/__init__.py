from .map_builder import SPICEComposedMapBuilder, ComposedMapBuilder  # noqa: F401
