// Default scene of the reference app: a sphere-grid fractal carved out of a bounding sphere.
// Restated from /root/reference/client/public/examples/guide.glsl (uniform annotations :2-48,
// material functions :51-88, sdf :91-102).  The //@key=value comments are the parameter
// annotations parsed by raymarching_engine_b200.params (CustomShaderParamParser.tsx:91-165).
uniform float bigSphereSize;
//@name="Big Sphere Size" @min=0 @step=0.001 @sensitivity=0.001 @default=4 @scale=linear
//@tooltip="Size of the big sphere that bounds the fractal." @format=numerical

uniform vec3 fractalColor;
//@name="Fractal Color" @tooltip="Diffuse color of the fractal." @default=0.5,0.5,0.5 @format=color

uniform float fractalIterations;
//@name="Fractal Iterations" @min=0 @max=20 @step=1 @sensitivity=0.01 @default=8
//@tooltip="Number of sphere grids in the fractal."

uniform float gridScaleFactor;
//@name="Grid Scale Factor" @min=0 @max=1 @step=0.001 @sensitivity=0.0003 @default=0.33333333333
//@tooltip="Factor by which successive sphere grids are scaled."

uniform vec3 bigSphereCenter;
//@name="Big Sphere Center" @step=0.001 @sensitivity=0.01 @default=0,0,10
//@tooltip="Center of the big sphere that bounds the fractal." @format=position/numerical

vec3 sceneDiffuseColor(vec3 position) {
  if (length(position) > 35.0) return vec3(0.0);
  return vec3(fractalColor);
}

vec3 sceneSpecularColor(vec3 position) {
  if (length(position) > 35.0) return vec3(0.0);
  return vec3(0.6);
}

float sceneSpecularRoughness(vec3 position) {
  return 0.2;
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
  float d = max(normalize(position).y, 0.2);
  vec3 brightColor = vec3(0.7, 0.8, 1.0) * d * 1.0;
  return (length(position) > 36.0) ? (brightColor * 2.00) : vec3(0.0);
}

float sdf(vec3 position) {
  float minDist = 9999.9;
  for (float i = -1.0; i < fractalIterations; i++) {
      float sf = pow(gridScaleFactor, i);
      vec3 d = abs(mod(position + vec3(0.5 * sf), sf)
         - vec3(sf / 2.0)) - vec3(sf / 3.0);
      float dist = length(d) - 0.21 * sf;
      minDist = min(dist, minDist);
  }
  minDist = max(length(position - bigSphereCenter) - bigSphereSize, -minDist);
  return minDist;
}
