"""remixfusion_b200 — B200-native (sm_100a) mapping hot path of RemixFusion: TSDF integration and the
mixed-representation ray query/render/backward, behind the reference's own Python entry points.
See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
