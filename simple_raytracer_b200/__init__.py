"""B200-native path-tracing kernel behind Simple-Raytracer's Tracer / render.cl contract."""
