// Infinite grid of mirror-like spheres with its own material set.
// Restated from /root/reference/client/dist/examples/sphere-grid.glsl:3-49 (the file exists only
// in the built bundle although the UI lists it, ShaderCodeSettings.tsx:83).
vec3 sceneDiffuseColor(vec3 position) {
  if (length(position) > 35.0) return vec3(0.0);
  return vec3(0.5);
}

vec3 sceneSpecularColor(vec3 position) {
  if (length(position) > 35.0) return vec3(0.0);
  return vec3(0.9);
}

float sceneSpecularRoughness(vec3 position) {
  return 0.01;
}

float sceneSubsurfaceScattering(vec3 position) {
  return 11111115.0;
}

vec3 sceneSubsurfaceScatteringColor(vec3 position) {
  if (length(position) > 30.0) return vec3(1.0);
  return vec3(1.0);
}

float sceneIOR(vec3 position) {
  return 100.0;
}

vec3 sceneEmission(vec3 position) {
  float d = max(normalize(position).x, 0.0);
  vec3 brightColor = vec3(0.7, 0.8, 1.0) * d * 1.0;
  return (length(position) > 36.0) ? (brightColor * 2.00) : vec3(0.0);
}

float sd_sphere(vec3 p, float radius, vec3 position) {
  return length(p - position) - radius;
}

float sdf(vec3 p) {
  vec3 repeat = mod(p + 1.0, vec3(2.0)) - 1.0;
  return sd_sphere(repeat, 0.4, vec3(0.0));
}
