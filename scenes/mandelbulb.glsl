// Synthetic divergence-stress scene (BASELINE.json config 4; NOT part of the reference):
// power-8 Mandelbulb distance estimator with an in-DE bailout, so the trip count of the inner
// loop depends on the sample position.  Frozen here; oracle twin: oracle/oracle.cpp SceneMandelbulb.
uniform float power;
//@name="Power" @min=2 @max=16 @step=1 @default=8

uniform float bailout;
//@name="Bailout radius" @min=1 @max=8 @step=0.01 @default=2

uniform float maxIterations;
//@name="DE iterations" @min=1 @max=32 @step=1 @default=12

float sdf(vec3 position) {
  vec3 z = position;
  float dr = 1.0;
  float r = 0.0;
  for (float i = 0.0; i < maxIterations; i++) {
    r = length(z);
    if (r > bailout) break;
    float theta = acos(z.z / r);
    float phi = atan(z.y, z.x);
    dr = pow(r, power - 1.0) * power * dr + 1.0;
    float zr = pow(r, power);
    theta = theta * power;
    phi = phi * power;
    z = zr * vec3(sin(theta) * cos(phi), sin(phi) * sin(theta), cos(theta));
    z += position;
  }
  return 0.5 * log(r) * r / dr;
}
