// drift-budget ablation: physics only, per-quantity precision switches
#include <cmath>
#include <cstdint>
#include <cstring>
extern "C" {
// state per rocket: double[10]: x,y,ang,dx,dy,dang,dry,fuel,kprev,fresh ; out: vx,vy,om
static inline float sinpoly(float t,float u){return fmaf(t*u, fmaf(fmaf(-0.0001958794891834259f,u,0.008332748897373676f),u,-0.166666641831398f),t);}
static inline float cospoly(float u){return fmaf(u, fmaf(fmaf(fmaf(2.446382313792128e-05f,u,-0.001388759003020823f),u,0.04166664928197861f),u,-0.5f),1.0f);}
static void sincos_deg(float a,float&s,float&c){
  float q=rintf(a*(1.0f/90.0f)); float r=fmaf(q,-90.0f,a); float t=r*0.017453292519943295f; float u=t*t;
  float sn=sinpoly(t,u), cs=cospoly(u); int qi=(int)q; bool sw=qi&1; float ss=sw?cs:sn, cc=sw?sn:cs;
  if(qi&2) ss=-ss; if((qi+1)&2) cc=-cc; s=ss;c=cc;}
static double norm_angle(double a,int&k){k=0; if(!(a>=-180.0&&a<180.0)){double t=floor((a+180.0)/360.0); a-=360.0*t;k=(int)t; if(a>=180){a-=360;++k;} if(a<-180){a+=360;--k;}} return a;}
#define RND(bit,v) ((bits&(bit))?(double)(v):(double)(float)(v))
void abl_step(int bits,double*st,int64_t n,const float*act,double*out){
  const double dtd=0.1; const float dt=0.1f,inv_dt=10.0f,dt2=(float)(0.01),half_dt2=(float)0.005;
  const float g=-9.81f, TP=1e7f, dragk=(float)(0.5*1.225*0.8*10.8), alphak=(float)(7.0e4*1.85/(0.5*1.85*1.85)*(180.0/M_PI)), damp=(float)0.995, burn=170.0f;
  for(int64_t i=0;i<n;++i){ double*s=st+10*i;
    double x=s[0],y=s[1],ang=s[2],sx=s[3],sy=s[4],sang=s[5],dry=s[6],fuel=s[7]; int kprev=(int)s[8]; bool fresh=s[9]!=0;
    float thr=act[2*i],cgr=act[2*i+1];
    float th=(fuel<=0)?0.f:fminf(fmaxf(thr,0.f),1.f); float cg=fminf(fmaxf(cgr,-1.f),1.f);
    double ax,ay,alpha,a0x,a0y; double vxb,vyb,omb;
    double vs=fresh?1.0:10.0; vxb=sx*vs; vyb=sy*vs; omb=sang*vs;
    if(bits&128){ // PROPOSED: fuel fixed (2^-13 kg), theta fixed (360/2^32), dtheta double, alpha dt2 double, 1/m double->float
      double m=dry+fuel; float r0=1.0f/(float)m; double im_d=(double)r0*(2.0-m*(double)r0); float im=(float)im_d;
      float vx=(float)vxb,vy=(float)vyb; float v2=fmaf(vx,vx,vy*vy); float kd=v2>1e-9f?dragk*sqrtf(v2)*im:0.f;
      float fa0x=-kd*vx, fa0y=fmaf(-kd,vy,g); float sn,cs; sincos_deg((float)ang,sn,cs);
      float t=th>1e-6f?th*TP*im:0.f; ax=fmaf(t,sn,fa0x); ay=fmaf(t,cs,fa0y); a0x=fa0x;a0y=fa0y;
      alpha=(double)cg*(7.0e4*1.85/(0.5*1.85*1.85)*(180.0/M_PI))*im_d;
    } else if(bits&16){ // forces in double
      double m=dry+fuel; double im=1.0/m; double v2=vxb*vxb+vyb*vyb; double kd=v2>1e-9? (0.5*1.225*0.8*10.8)*sqrt(v2)*im:0.0;
      a0x=-kd*vxb; a0y=-kd*vyb-9.81; double rad=ang*M_PI/180.0; double t=th>1e-6?(double)th*1e7*im:0.0;
      ax=t*sin(rad)+a0x; ay=t*cos(rad)+a0y; alpha=(double)cg*(7.0e4*1.85/(0.5*1.85*1.85)*(180.0/M_PI))*im;
    } else {
      float m,im; if(bits&32){double md=dry+fuel; im=(float)(1.0/md);} else {m=(float)dry+(float)fuel; im=1.0f/m;}
      float vx=(float)vxb,vy=(float)vyb; float v2=fmaf(vx,vx,vy*vy); float kd=v2>1e-9f?dragk*sqrtf(v2)*im:0.f;
      float fa0x=-kd*vx, fa0y=fmaf(-kd,vy,g); float sn,cs; 
      if(bits&64){ double rad=ang*M_PI/180.0; sn=(float)sin(rad); cs=(float)cos(rad);} else sincos_deg((float)ang,sn,cs);
      float t=th>1e-6f?th*TP*im:0.f; ax=fmaf(t,sn,fa0x); ay=fmaf(t,cs,fa0y); alpha=cg*alphak*im; a0x=fa0x;a0y=fa0y;
    }
    double dx,dy,dang; 
    if(fresh){ if(bits&4) dx=vxb*dtd-0.5*a0x*dtd*dtd; else dx=fmaf(-half_dt2,(float)a0x,(float)vxb*dt);
               if(bits&(8|256)) dy=vyb*dtd-0.5*a0y*dtd*dtd; else dy=fmaf(-half_dt2,(float)a0y,(float)vyb*dt);
               double back=(bits&2)?omb*dtd:(double)((float)omb*dt); int kp; norm_angle(ang-back,kp); dang=back+360.0*kp; }
    else { dx=sx;dy=sy;dang=sang-360.0*kprev; }
    double ndx,ndy,ndang; const double XU=1.0/8388608.0;
    if(bits&1024){ if(fresh){dx=rint(dx/XU)*XU; dy=rint(dy/XU)*XU;} }
    if(bits&1024) ndx=dx+rint((double)((float)ax*dt2)/XU)*XU; else if(bits&4) ndx=dx+ax*(dtd*dtd); else ndx=fmaf((float)ax,dt2,(float)dx);
    if(bits&1024) ndy=dy+rint((double)((float)ay*dt2)/XU)*XU; else if(bits&(8|256)) ndy=dy+ay*(dtd*dtd); else ndy=fmaf((float)ay,dt2,(float)dy);
    if(bits&2) ndang=dang*0.995+alpha*(dtd*dtd); else ndang=fmaf((float)dang,damp,(float)alpha*dt2);
    if(bits&2048){ float hi=(float)ndang; double r=ndang-(double)hi; int e; frexp((double)hi,&e); double u=ldexp(1.0,e-24-8); double lo=rint(r/u); if(lo>127)lo=127; if(lo<-128)lo=-128; ndang=(double)hi+lo*u; }
    double nx=(bits&4096)?(x+floor(ndx*32768.0+0.5)/32768.0):(bits&1024)?(double)((float)x+(float)ndx):(bits&4)?x+ndx:(double)((float)x+(float)ndx);
    double ny=(bits&4096)?(y+floor(ndy*32768.0+0.5)/32768.0):(bits&1024)?(double)((float)y+(float)ndy):(bits&(8|512))?y+ndy:(double)((float)y+(float)ndy);
    int k; double nang; if(bits&128){ const double U=360.0/4294967296.0; nang=norm_angle(ang+ndang,k); nang=rint(nang/U)*U; if(nang>=180.0){nang-=360.0;++k;} } else nang=(bits&2)?norm_angle(ang+ndang,k):(double)(float)norm_angle((double)((float)ang+(float)ndang),k);
    double nfuel; if(bits&128){ const double FU=1.0/8192.0; nfuel=fmax(0.0,fuel-rint(fmax(0.0,(double)th*170.0)/FU)*FU);} else nfuel=(bits&1)?fmax(0.0,fuel-fmax(0.0,(double)th*170.0)):(double)fmaxf(0.f,(float)fuel-fmaxf(0.f,th*burn));
    s[0]=nx;s[1]=ny;s[2]=nang;s[3]=ndx;s[4]=ndy;s[5]=ndang;s[7]=nfuel;s[8]=k;s[9]=0;
    out[3*i]=ndx*10.0; out[3*i+1]=ndy*10.0; out[3*i+2]=ndang*10.0;
  }
}
}
