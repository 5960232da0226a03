// Placeholder scene the reference app starts with before guide.glsl is fetched.
// Restated from /root/reference/client/src/index.tsx:374-388 (sdf only; the material functions at
// :337-372 equal the injected defaults).
float sdf(vec3 position) {
  float minDist = 9999.9;
  for (float i = -1.0; i < 10.0; i++) {
      float sf = pow(0.3333333333333, i);
      ivec3 index = ivec3(position / 2.0);
      vec3 d = abs(mod(position + vec3(0.5 * sf), sf) - vec3(sf / 2.0)) - vec3(sf / 3.0);
      float dist = length(d) - 0.21 * sf;
      minDist = min(dist, minDist);
  }
  minDist = max(length(position) - 5.0, -minDist);
  return minDist;
}
